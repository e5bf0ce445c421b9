// Minimal stand-in for <sl/clock.hpp> and the sl string/path utilities the
// reference's svbuilder prints with (see cstdint.hpp header note).
#pragma once
#include <chrono>
#include <sl/cstdint.hpp>
namespace sl {
class time_duration {
public:
	int64_t us_;
	time_duration() : us_(0) {}
	explicit time_duration(int64_t us) : us_(us) {}
	uint64_t as_microseconds() const { return uint64_t(us_); }
	uint64_t as_milliseconds() const { return uint64_t(us_ / 1000); }
	double as_seconds() const { return double(us_) * 1e-6; }
	time_duration operator+(const time_duration& o) const { return time_duration(us_ + o.us_); }
	time_duration operator-(const time_duration& o) const { return time_duration(us_ - o.us_); }
	time_duration& operator+=(const time_duration& o) { us_ += o.us_; return *this; }
	time_duration operator*(float f) const { return time_duration(int64_t(double(us_) * double(f))); }
};
class time_point {
public:
	int64_t us_;
	time_point() : us_(0) {}
	explicit time_point(int64_t us) : us_(us) {}
	time_duration operator-(const time_point& o) const { return time_duration(us_ - o.us_); }
};
class real_time_clock {
	int64_t t0_;
public:
	real_time_clock() { restart(); }
	static time_point now() {
		return time_point(std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count());
	}
	void restart() { t0_ = now().us_; }
	time_duration elapsed() const { return time_duration(now().us_ - t0_); }
};
inline std::string human_readable_duration(const time_duration& d) {
	char buf[64]; double s = d.as_seconds();
	if (s < 1e-3) snprintf(buf, sizeof buf, "%.0f us", s * 1e6);
	else if (s < 1.0) snprintf(buf, sizeof buf, "%.2f ms", s * 1e3);
	else if (s < 60.0) snprintf(buf, sizeof buf, "%.2f s", s);
	else snprintf(buf, sizeof buf, "%dm %.1fs", int(s / 60.0), s - 60.0 * int(s / 60.0));
	return buf;
}
inline std::string human_readable_quantity(double q) {
	char buf[64];
	if (q < 1e3) snprintf(buf, sizeof buf, "%.0f", q);
	else if (q < 1e6) snprintf(buf, sizeof buf, "%.2f K", q / 1e3);
	else if (q < 1e9) snprintf(buf, sizeof buf, "%.2f M", q / 1e6);
	else snprintf(buf, sizeof buf, "%.2f G", q / 1e9);
	return buf;
}
inline std::string human_readable_size(double q) {
	char buf[64];
	if (q < 1024.0) snprintf(buf, sizeof buf, "%.0f B", q);
	else if (q < 1048576.0) snprintf(buf, sizeof buf, "%.2f KB", q / 1024.0);
	else if (q < 1073741824.0) snprintf(buf, sizeof buf, "%.2f MB", q / 1048576.0);
	else snprintf(buf, sizeof buf, "%.2f GB", q / 1073741824.0);
	return buf;
}
inline std::string pathname_directory_separators() { return "/"; }
inline std::string pathname_directory(const std::string& p) {
	std::size_t k = p.find_last_of('/');
	return (k == std::string::npos) ? std::string(".") : p.substr(0, k);
}
inline std::string pathname_base(const std::string& p) {
	std::size_t k = p.find_last_of('/');
	return (k == std::string::npos) ? p : p.substr(k + 1);
}
inline std::string pathname_without_extension(const std::string& p) {
	std::size_t k = p.find_last_of('.'); std::size_t s = p.find_last_of('/');
	if (k == std::string::npos || (s != std::string::npos && k < s)) return p;
	return p.substr(0, k);
}
}
