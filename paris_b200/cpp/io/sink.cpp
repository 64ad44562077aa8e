// io/sink.cpp -- see sink.h.
#include "sink.h"

#include <mutex>
#include <stdexcept>
#include <system_error>
#include <utility>

#include "filesystem.h"
#include "log.h"

namespace paris
{
    namespace
    {
        // Runs `body`; failures of the file system or the container are logged with the reference's wording and
        // re-thrown as the pipeline's own exception types: system errors always become stage_runtime_error, other
        // runtime errors become `Other` (construction errors while the sink is built, runtime errors while saving).
        template <class Other, class Body>
        auto translating(const char* where, const char* failed, Body&& body) -> void
        {
            try
            {
                std::forward<Body>(body)();
            }
            catch(const stage_construction_error&)
            {
                throw;
            }
            catch(const std::system_error& se)
            {
                log::fatal() << where << ": system error: " << se.code() << " - " << se.what();
                throw stage_runtime_error{failed};
            }
            catch(const std::runtime_error& re)
            {
                log::fatal() << where << ": runtime error: " << re.what();
                throw Other{failed};
            }
        }
    }

    sink::sink(const std::string& path, const std::string& prefix, const volume_geometry& vol_geo)
    : path_{path}, vol_geo_(vol_geo)
    {
        translating<stage_construction_error>("sink::sink()", "sink::sink() failed", [&] {
            if(!create_directory(path))
            {
                log::fatal() << "sink::sink() failed to create output directory at " << path;
                throw stage_construction_error{"sink::sink() failed"};
            }
            // <path>/<prefix>; ddbvf::create appends the suffix
            if(path_.empty() || path_.back() != '/')
                path_.push_back('/');
            path_.append(prefix);
            handle_ = ddbvf::create(path_, vol_geo_.dim_x, vol_geo_.dim_y, vol_geo_.dim_z);
        });
    }

    auto sink::save(const b200::volume_device_type& v) -> void
    {
        save(v, v.off);
    }

    namespace
    {
        auto file_mutex() -> std::mutex&
        {
            static std::mutex m;
            return m;
        }
    }

    auto sink::save(const float* h_slab, std::uint32_t dim_x, std::uint32_t dim_y, std::uint32_t dim_z,
                    std::uint32_t first_slice) -> void
    {
        translating<stage_runtime_error>("sink::save()", "sink::save() failed", [&] {
            const std::lock_guard<std::mutex> lock{file_mutex()};
            ddbvf::write(handle_, h_slab, dim_x, dim_y, dim_z, first_slice);
        });
    }

    auto sink::save(const b200::volume_device_type& v, std::uint32_t first_slice) -> void
    {
        translating<stage_runtime_error>("sink::save()", "sink::save() failed", [&] {
            // entered concurrently by every device thread: the pinned staging volume and the download are per
            // thread, only the file is shared (src/sink.cpp:76-81)
            auto staged = b200::make_volume_host(v.dim_x, v.dim_y, v.dim_z);
            b200::copy_d2h(v, staged);

            const std::lock_guard<std::mutex> lock{file_mutex()};
            ddbvf::write(handle_, staged, first_slice);
        });
    }
}
