// Minimal stand-in for <sl/fixed_size_vector.hpp> (see cstdint.hpp header note).
// Arithmetic conventions chosen here (SL's own are not in the reference tree, so
// parity w.r.t. the real SL is UNPINNED):
//   dot   = a0*b0 + a1*b1 + a2*b2   evaluated left to right
//   cross = (a1*b2 - a2*b1, a2*b0 - a0*b2, a0*b1 - a1*b0)
#pragma once
#include <sl/cstdint.hpp>
#include <sl/clock.hpp>   // real SL pulls its string/path helpers in transitively
namespace sl {
template <std::size_t N, class T> class fixed_size_vector {
public:
	T v_[N];
	fixed_size_vector() { for (std::size_t i = 0; i < N; ++i) v_[i] = T(0); }
	fixed_size_vector(T a, T b, T c) { static_assert(N == 3, "3D only"); v_[0] = a; v_[1] = b; v_[2] = c; }
	T& operator[](std::size_t i) { return v_[i]; }
	const T& operator[](std::size_t i) const { return v_[i]; }
	T* to_pointer() { return v_; }
	const T* to_pointer() const { return v_; }
	T dot(const fixed_size_vector& o) const {
		T s = v_[0] * o.v_[0];
		for (std::size_t i = 1; i < N; ++i) s = s + v_[i] * o.v_[i];
		return s;
	}
	fixed_size_vector cross(const fixed_size_vector& o) const {
		return fixed_size_vector(v_[1] * o.v_[2] - v_[2] * o.v_[1],
		                         v_[2] * o.v_[0] - v_[0] * o.v_[2],
		                         v_[0] * o.v_[1] - v_[1] * o.v_[0]);
	}
	T two_norm() const { return T(std::sqrt(double(dot(*this)))); }
	fixed_size_vector ok_normalized() const {
		T n = two_norm(); fixed_size_vector r;
		if (n > T(0)) for (std::size_t i = 0; i < N; ++i) r.v_[i] = v_[i] / n;
		return r;
	}
	fixed_size_vector operator+(const fixed_size_vector& o) const { fixed_size_vector r; for (std::size_t i = 0; i < N; ++i) r.v_[i] = v_[i] + o.v_[i]; return r; }
	fixed_size_vector operator-(const fixed_size_vector& o) const { fixed_size_vector r; for (std::size_t i = 0; i < N; ++i) r.v_[i] = v_[i] - o.v_[i]; return r; }
	fixed_size_vector operator-() const { fixed_size_vector r; for (std::size_t i = 0; i < N; ++i) r.v_[i] = -v_[i]; return r; }
	fixed_size_vector operator*(T s) const { fixed_size_vector r; for (std::size_t i = 0; i < N; ++i) r.v_[i] = v_[i] * s; return r; }
	fixed_size_vector operator/(T s) const { fixed_size_vector r; for (std::size_t i = 0; i < N; ++i) r.v_[i] = v_[i] / s; return r; }
	fixed_size_vector& operator+=(const fixed_size_vector& o) { for (std::size_t i = 0; i < N; ++i) v_[i] += o.v_[i]; return *this; }
	bool operator==(const fixed_size_vector& o) const { for (std::size_t i = 0; i < N; ++i) if (!(v_[i] == o.v_[i])) return false; return true; }
};
template <std::size_t N, class T> inline fixed_size_vector<N, T> operator*(T s, const fixed_size_vector<N, T>& v) { return v * s; }
template <std::size_t N, class T> inline std::ostream& operator<<(std::ostream& os, const fixed_size_vector<N, T>& v) {
	for (std::size_t i = 0; i < N; ++i) os << (i ? " " : "") << v[i];
	return os;
}
typedef fixed_size_vector<3, float>  vector3f;
typedef fixed_size_vector<3, double> vector3d;
typedef fixed_size_vector<3, float>  color3f;
}
