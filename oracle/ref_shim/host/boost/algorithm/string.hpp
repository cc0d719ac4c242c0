/*
 * boost/algorithm/string.hpp stand-in: split / is_any_of / trim with Boost's semantics -- split keeps empty
 * tokens (token_compress_off), so "Kd  0.3" yields an empty field, which the reference's parsers then read
 * the way they do upstream.  TEST INFRASTRUCTURE.
 */
#ifndef PBR_REF_BOOST_STRING_HPP
#define PBR_REF_BOOST_STRING_HPP

#include <ctype.h>
#include <string>
#include <vector>

namespace boost {

struct is_any_of_pred {
	std::string chars;
	bool operator()(char c) const { return chars.find(c) != std::string::npos; }
};

inline is_any_of_pred is_any_of(const char* chars) { is_any_of_pred p; p.chars = chars; return p; }

template <typename Pred>
inline std::vector<std::string>& split(std::vector<std::string>& out, const std::string& in, Pred pred) {
	out.clear();
	std::string cur;
	for (size_t i = 0; i < in.size(); i++) {
		if (pred(in[i])) { out.push_back(cur); cur.clear(); }
		else cur.push_back(in[i]);
	}
	out.push_back(cur);
	return out;
}

namespace algorithm {

inline void trim(std::string& s) {
	size_t a = 0, b = s.size();
	while (a < b && isspace((unsigned char) s[a])) a++;
	while (b > a && isspace((unsigned char) s[b - 1])) b--;
	s = s.substr(a, b - a);
}

using boost::split;
using boost::is_any_of;

} /* namespace algorithm */

using algorithm::trim;

} /* namespace boost */

#endif
