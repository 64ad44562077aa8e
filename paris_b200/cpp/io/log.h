// io/log.h -- severity-prefixed stderr logging in place of Boost.Log's BOOST_LOG_TRIVIAL (used by the
// reference for diagnostics only; nothing on the data path depends on it).
#pragma once

#include <iostream>
#include <mutex>
#include <sstream>

namespace paris
{
    namespace log
    {
        enum class severity { debug, info, warning, error, fatal };

        inline auto threshold() noexcept -> severity&
        {
            static severity s = severity::info;   // (the reference's release filter, src/main.cpp:63)
            return s;
        }

        class line
        {
            public:
                explicit line(severity s) : active_{s >= threshold()}
                {
                    static const char* const names[] = {"debug", "info", "warning", "error", "fatal"};
                    if(active_)
                        text_ << '[' << names[static_cast<int>(s)] << "] ";
                }
                line(line&&) = default;
                ~line()
                {
                    if(!active_)
                        return;
                    static std::mutex m;
                    std::lock_guard<std::mutex> lock{m};
                    std::cerr << text_.str() << std::endl;
                }
                template <class T>
                auto operator<<(const T& v) -> line&
                {
                    if(active_)
                        text_ << v;
                    return *this;
                }

            private:
                bool active_;
                std::ostringstream text_;
        };

        inline auto debug() -> line { return line{severity::debug}; }
        inline auto info() -> line { return line{severity::info}; }
        inline auto warning() -> line { return line{severity::warning}; }
        inline auto error() -> line { return line{severity::error}; }
        inline auto fatal() -> line { return line{severity::fatal}; }
    }
}
