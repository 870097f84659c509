// include-path shim: lets a MEmilio program keep `#include "ode_secirvvs/model.h"` when built against the B200 facade
// (add -Iinclude/memilio_b200/compat). The model lives in include/memilio_b200/osecirvvs.hpp.
#pragma once
#include "../../mio.hpp"
