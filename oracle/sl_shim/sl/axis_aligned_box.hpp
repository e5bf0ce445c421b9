// Minimal stand-in for <sl/axis_aligned_box.hpp> (see cstdint.hpp header note).
// Conventions chosen here (UNPINNED w.r.t. the real SL):
//   center()            = lo + (hi - lo) * 0.5      -> see note below
//   half_side_lengths() = (hi - lo) * 0.5
// center() is written as (lo + hi) * 0.5 component-wise.
#pragma once
#include <sl/fixed_size_point.hpp>
namespace sl {
template <std::size_t N, class T> class axis_aligned_box {
public:
	fixed_size_point<N, T> p_[2];
	axis_aligned_box() { to_empty(); }
	axis_aligned_box(const fixed_size_point<N, T>& lo, const fixed_size_point<N, T>& hi) { p_[0] = lo; p_[1] = hi; }
	fixed_size_point<N, T>& operator[](std::size_t i) { return p_[i]; }
	const fixed_size_point<N, T>& operator[](std::size_t i) const { return p_[i]; }
	void to_empty() {
		for (std::size_t i = 0; i < N; ++i) { p_[0][i] = std::numeric_limits<T>::max(); p_[1][i] = -std::numeric_limits<T>::max(); }
	}
	void merge(const fixed_size_point<N, T>& p) {
		for (std::size_t i = 0; i < N; ++i) { if (p[i] < p_[0][i]) p_[0][i] = p[i]; if (p[i] > p_[1][i]) p_[1][i] = p[i]; }
	}
	fixed_size_point<N, T> center() const {
		fixed_size_point<N, T> c;
		for (std::size_t i = 0; i < N; ++i) c[i] = (p_[0][i] + p_[1][i]) * T(0.5);
		return c;
	}
	fixed_size_vector<N, T> half_side_lengths() const {
		fixed_size_vector<N, T> h;
		for (std::size_t i = 0; i < N; ++i) h[i] = (p_[1][i] - p_[0][i]) * T(0.5);
		return h;
	}
	bool contains(const fixed_size_point<N, T>& p) const {
		for (std::size_t i = 0; i < N; ++i) if (p[i] < p_[0][i] || p[i] > p_[1][i]) return false;
		return true;
	}
};
template <std::size_t N, class T> inline std::ostream& operator<<(std::ostream& os, const axis_aligned_box<N, T>& b) {
	return os << "[" << b[0] << "] [" << b[1] << "]";
}
typedef axis_aligned_box<3, float>  aabox3f;
typedef axis_aligned_box<3, double> aabox3d;
template <class To> struct conv_to { template <class From> static To from(const From& b) { To r; for (int k = 0; k < 2; ++k) for (int i = 0; i < 3; ++i) r[k][i] = b[k][i]; return r; } };
}
