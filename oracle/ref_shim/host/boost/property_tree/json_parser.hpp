/*
 * boost::property_tree::json_parser::read_json stand-in: flattens a JSON object of objects / numbers /
 * strings / booleans into the dotted keys of the ptree stand-in.  TEST INFRASTRUCTURE.
 */
#ifndef PBR_REF_BOOST_JSON_PARSER_HPP
#define PBR_REF_BOOST_JSON_PARSER_HPP

#include <ctype.h>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>

#include "ptree.hpp"

namespace boost { namespace property_tree { namespace json_parser {

namespace detail {

inline void skipSpace(const std::string& s, size_t& i) {
	while (i < s.size() && isspace((unsigned char) s[i])) i++;
}

inline std::string parseString(const std::string& s, size_t& i) {
	std::string out;
	i++;                                       /* opening quote */
	while (i < s.size() && s[i] != '"') {
		if (s[i] == '\\' && i + 1 < s.size()) i++;
		out.push_back(s[i++]);
	}
	i++;
	return out;
}

inline void parseValue(const std::string& s, size_t& i, const std::string& prefix, ptree& tree) {
	skipSpace(s, i);
	if (i >= s.size()) throw std::runtime_error("read_json: unexpected end");
	if (s[i] == '{') {
		i++;
		while (true) {
			skipSpace(s, i);
			if (s[i] == '}') { i++; return; }
			const std::string key = parseString(s, i);
			skipSpace(s, i);
			if (s[i] != ':') throw std::runtime_error("read_json: expected ':'");
			i++;
			parseValue(s, i, prefix.empty() ? key : prefix + "." + key, tree);
			skipSpace(s, i);
			if (s[i] == ',') i++;
		}
	}
	else if (s[i] == '"') {
		tree.values[prefix] = parseString(s, i);
	}
	else {
		std::string tok;
		while (i < s.size() && s[i] != ',' && s[i] != '}' && !isspace((unsigned char) s[i])) tok.push_back(s[i++]);
		tree.values[prefix] = tok;
	}
}

} /* namespace detail */

inline void read_json(const std::string& filename, ptree& tree) {
	std::ifstream in(filename.c_str());
	if (!in.good()) throw std::runtime_error("read_json: cannot open " + filename);
	std::stringstream buf;
	buf << in.rdbuf();
	const std::string s = buf.str();
	size_t i = 0;
	detail::parseValue(s, i, "", tree);
}

} } }

#endif
