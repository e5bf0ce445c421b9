// Minimal stand-in for <sl/fixed_size_point.hpp> (see cstdint.hpp header note).
#pragma once
#include <sl/fixed_size_vector.hpp>
namespace sl {
template <std::size_t N, class T> class fixed_size_point {
public:
	T v_[N];
	fixed_size_point() { for (std::size_t i = 0; i < N; ++i) v_[i] = T(0); }
	fixed_size_point(T a, T b, T c) { static_assert(N == 3, "3D only"); v_[0] = a; v_[1] = b; v_[2] = c; }
	T& operator[](std::size_t i) { return v_[i]; }
	const T& operator[](std::size_t i) const { return v_[i]; }
	T* to_pointer() { return v_; }
	const T* to_pointer() const { return v_; }
	fixed_size_vector<N, T> as_vector() const { fixed_size_vector<N, T> r; for (std::size_t i = 0; i < N; ++i) r[i] = v_[i]; return r; }
	fixed_size_point operator+(const fixed_size_vector<N, T>& o) const { fixed_size_point r; for (std::size_t i = 0; i < N; ++i) r.v_[i] = v_[i] + o[i]; return r; }
	fixed_size_point operator-(const fixed_size_vector<N, T>& o) const { fixed_size_point r; for (std::size_t i = 0; i < N; ++i) r.v_[i] = v_[i] - o[i]; return r; }
	fixed_size_vector<N, T> operator-(const fixed_size_point& o) const { fixed_size_vector<N, T> r; for (std::size_t i = 0; i < N; ++i) r[i] = v_[i] - o.v_[i]; return r; }
	fixed_size_point& operator+=(const fixed_size_vector<N, T>& o) { for (std::size_t i = 0; i < N; ++i) v_[i] += o[i]; return *this; }
	bool operator==(const fixed_size_point& o) const { for (std::size_t i = 0; i < N; ++i) if (!(v_[i] == o.v_[i])) return false; return true; }
	T distance_to(const fixed_size_point& o) const { return (*this - o).two_norm(); }
};
template <std::size_t N, class T> inline std::ostream& operator<<(std::ostream& os, const fixed_size_point<N, T>& v) {
	for (std::size_t i = 0; i < N; ++i) os << (i ? " " : "") << v[i];
	return os;
}
typedef fixed_size_point<3, float>  point3f;
typedef fixed_size_point<3, double> point3d;
}
