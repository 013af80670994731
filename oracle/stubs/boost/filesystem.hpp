// Minimal stand-in for Boost.Filesystem (not installed in this image).
// TEST INFRASTRUCTURE ONLY (oracle/_ref): provides the names main.h / APD.h mention.
#ifndef DVP_ORACLE_BOOST_FS_STUB_HPP
#define DVP_ORACLE_BOOST_FS_STUB_HPP
#include <string>
#include <fstream>
#include <ostream>
namespace boost { namespace filesystem {
class path {
	std::string s_;
public:
	path() {}
	path(const char* s) : s_(s) {}
	path(const std::string& s) : s_(s) {}
	const std::string& string() const { return s_; }
	path operator/(const path& o) const { return path(s_ + "/" + o.s_); }
	operator std::string() const { return s_; }   // boost's ifstream / ofstream open a path; std::ifstream opens a string
	friend std::ostream& operator<<(std::ostream& os, const path& p) { return os << p.s_; }
};
inline bool create_directory(const path&) { return true; }
inline bool exists(const path& p) { std::ifstream f(p.string().c_str()); return f.good(); }   // InuputInitialization checks weak.bin (APD.cpp:1171)   // GenerateSampleList (main.cpp:152) makes its result folders; the checker does not
typedef std::ifstream ifstream;
typedef std::ofstream ofstream;
} }
#endif
