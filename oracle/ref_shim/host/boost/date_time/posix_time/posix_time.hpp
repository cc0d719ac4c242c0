/* boost::posix_time stand-in for the reference's timers (they only feed log lines).  TEST INFRASTRUCTURE. */
#ifndef PBR_REF_BOOST_POSIX_TIME_HPP
#define PBR_REF_BOOST_POSIX_TIME_HPP

#include <chrono>

namespace boost { namespace posix_time {

struct time_duration {
	long long us;
	long long total_milliseconds() const { return us / 1000; }
	long long total_microseconds() const { return us; }
};

struct ptime {
	std::chrono::steady_clock::time_point t;
};

inline time_duration operator-(const ptime& a, const ptime& b) {
	time_duration d;
	d.us = std::chrono::duration_cast<std::chrono::microseconds>(a.t - b.t).count();
	return d;
}

struct microsec_clock {
	static ptime local_time() { ptime p; p.t = std::chrono::steady_clock::now(); return p; }
};

} }

#endif
