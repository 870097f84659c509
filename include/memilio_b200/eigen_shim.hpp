// eigen_shim.hpp -- the small subset of Eigen::VectorX<T> / MatrixX<T> that MEmilio's Graph-ODE examples touch
// (Constant, Zero, operator[], operator(), size, rows, cols, sum, array-free arithmetic). Skipped when the real Eigen
// has been included first.
#pragma once

#ifndef EIGEN_CORE_H
#define MIO_B200_EIGEN_SHIM 1

#include <cstddef>
#include <initializer_list>
#include <vector>

namespace Eigen
{
using Index = std::ptrdiff_t;

template <class T>
class VectorX
{
public:
    VectorX() = default;
    explicit VectorX(Index n)
        : m_v(size_t(n))
    {
    }
    static VectorX Constant(Index n, T v)
    {
        VectorX r(n);
        for (auto& x : r.m_v) {
            x = v;
        }
        return r;
    }
    static VectorX Zero(Index n)
    {
        return Constant(n, T(0));
    }
    Index size() const
    {
        return Index(m_v.size());
    }
    Index rows() const
    {
        return size();
    }
    T& operator[](Index i)
    {
        return m_v[size_t(i)];
    }
    const T& operator[](Index i) const
    {
        return m_v[size_t(i)];
    }
    T& operator()(Index i)
    {
        return m_v[size_t(i)];
    }
    const T& operator()(Index i) const
    {
        return m_v[size_t(i)];
    }
    T sum() const
    {
        T s = T(0);
        for (auto& x : m_v) {
            s += x;
        }
        return s;
    }
    T* data()
    {
        return m_v.data();
    }
    const T* data() const
    {
        return m_v.data();
    }
    VectorX& operator+=(const VectorX& o)
    {
        for (size_t i = 0; i < m_v.size(); ++i) {
            m_v[i] += o.m_v[i];
        }
        return *this;
    }
    VectorX& operator-=(const VectorX& o)
    {
        for (size_t i = 0; i < m_v.size(); ++i) {
            m_v[i] -= o.m_v[i];
        }
        return *this;
    }
    friend VectorX operator/(VectorX a, T d)
    {
        for (auto& x : a.m_v) {
            x /= d;
        }
        return a;
    }
    friend VectorX operator*(T d, VectorX a)
    {
        for (auto& x : a.m_v) {
            x *= d;
        }
        return a;
    }
    auto begin() const
    {
        return m_v.begin();
    }
    auto end() const
    {
        return m_v.end();
    }

private:
    std::vector<T> m_v;
};

// row-major dense matrix
template <class T>
class MatrixX
{
public:
    MatrixX() = default;
    MatrixX(Index r, Index c)
        : m_r(r)
        , m_c(c)
        , m_v(size_t(r * c))
    {
    }
    static MatrixX Constant(Index r, Index c, T v)
    {
        MatrixX m(r, c);
        for (auto& x : m.m_v) {
            x = v;
        }
        return m;
    }
    static MatrixX Zero(Index r, Index c)
    {
        return Constant(r, c, T(0));
    }
    Index rows() const
    {
        return m_r;
    }
    Index cols() const
    {
        return m_c;
    }
    T& operator()(Index i, Index j)
    {
        return m_v[size_t(i * m_c + j)];
    }
    const T& operator()(Index i, Index j) const
    {
        return m_v[size_t(i * m_c + j)];
    }
    const T* data() const
    {
        return m_v.data();
    }
    T* data()
    {
        return m_v.data();
    }
    Index size() const
    {
        return m_r * m_c;
    }
    bool operator==(const MatrixX& o) const
    {
        return m_r == o.m_r && m_c == o.m_c && m_v == o.m_v;
    }
    void setConstant(T v)
    {
        for (auto& x : m_v) {
            x = v;
        }
    }
    // m.diagonal().setConstant(v)
    struct DiagonalView {
        MatrixX& m;
        void setConstant(T v)
        {
            for (Index i = 0; i < (m.m_r < m.m_c ? m.m_r : m.m_c); ++i) {
                m(i, i) = v;
            }
        }
    };
    DiagonalView diagonal()
    {
        return DiagonalView{*this};
    }

private:
    Index m_r = 0, m_c = 0;
    std::vector<T> m_v;
};

using VectorXd = VectorX<double>;
using MatrixXd = MatrixX<double>;
} // namespace Eigen

#endif // EIGEN_CORE_H
