// include-path shim: lets a MEmilio program keep `#include "memilio/compartments/flow_simulation_base.h"` when built against the B200 facade
// (add -Iinclude/memilio_b200/compat). Everything lives in include/memilio_b200/mio.hpp.
#pragma once
#include "../../../mio.hpp"
