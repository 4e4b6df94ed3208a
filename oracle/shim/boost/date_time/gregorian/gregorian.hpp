/* Test-infrastructure shim (oracle build only): src/core/configuration.cpp:118-132 prints a date
 * through boost::gregorian; only construction and to_simple_string are needed. */
#ifndef SHKZ_ORACLE_SHIM_BOOST_GREGORIAN_HPP
#define SHKZ_ORACLE_SHIM_BOOST_GREGORIAN_HPP
#include <cstdio>
#include <string>
namespace boost {
namespace gregorian {
class date {
public:
	date(int year, int month, int day) : m_year(year), m_month(month), m_day(day) {}
	std::string simple() const {
		static const char *names[] = {"Jan","Feb","Mar","Apr","May","Jun","Jul","Aug","Sep","Oct","Nov","Dec"};
		char buf[40];
		std::snprintf(buf, sizeof buf, "%04d-%s-%02d", m_year, names[(m_month + 11) % 12], m_day);
		return buf;
	}
private:
	int m_year, m_month, m_day;
};
inline std::string to_simple_string(const date &d) { return d.simple(); }
}
}
#endif
