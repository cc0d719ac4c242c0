/* boost::posix_time stand-in for the reference's timers.  They feed log lines -- and the per-frame seed
 * (PathTracer::getTimeSinceStart, PathTracer.cpp:78-82): clockOverrideUs() lets a test set the time.
 * TEST INFRASTRUCTURE. */
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
	long long us;
};

inline time_duration operator-(const ptime& a, const ptime& b) {
	time_duration d;
	d.us = a.us - b.us;
	return d;
}

/* >= 0: local_time() returns this many microseconds instead of the real clock */
inline long long& clockOverrideUs() { static long long v = -1; return v; }

struct microsec_clock {
	static ptime local_time() {
		ptime p;
		p.us = clockOverrideUs() >= 0 ? clockOverrideUs()
			: std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
		return p;
	}
};

} }

#endif
