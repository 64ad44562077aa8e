/*
 * oracle/shim/boost/log/trivial.hpp -- TEST INFRASTRUCTURE, not product code.
 *
 * Null-stream stand-in for Boost.Log so the reference sources that only log
 * (/root/reference/src/geometry.cpp:79-127, src/backprojection.cpp:66) compile
 * unmodified in an image without Boost.  Everything streamed is discarded.
 */
#ifndef PARIS_B200_ORACLE_BOOST_LOG_SHIM_HPP_
#define PARIS_B200_ORACLE_BOOST_LOG_SHIM_HPP_

#include <ios>

namespace oracle_shim
{
    struct null_stream
    {
        template <class T>
        const null_stream& operator<<(const T&) const noexcept { return *this; }
        // manipulators such as std::setprecision / std::endl
        const null_stream& operator<<(std::ios_base& (*)(std::ios_base&)) const noexcept { return *this; }
    };
}

#define BOOST_LOG_TRIVIAL(lvl) ::oracle_shim::null_stream{}

#endif
