/*
 * boost::property_tree::ptree stand-in: a flat "dotted.key" -> text map with get<T>() / put(), which is all
 * the reference's Cfg class asks of it.  TEST INFRASTRUCTURE.
 */
#ifndef PBR_REF_BOOST_PTREE_HPP
#define PBR_REF_BOOST_PTREE_HPP

#include <map>
#include <sstream>
#include <stdexcept>
#include <string>

namespace boost { namespace property_tree {

class ptree {
	public:
		template <typename T> T get(const std::string& key) const {
			std::map<std::string, std::string>::const_iterator it = values.find(key);
			if (it == values.end()) throw std::runtime_error("No such node (" + key + ")");
			return convert<T>(it->second);
		}
		template <typename T> void put(const std::string& key, const T& value) {
			std::ostringstream o;
			o << value;
			values[key] = o.str();
		}
		std::map<std::string, std::string> values;

	private:
		template <typename T> static T convert(const std::string& text) {
			std::istringstream in(text);
			T v = T();
			in >> v;
			if (in.fail()) {                         /* "true" / "false" */
				std::istringstream in2(text);
				in2 >> std::boolalpha >> v;
			}
			return v;
		}
};

template <> inline std::string ptree::convert<std::string>(const std::string& text) { return text; }

} }

#endif
