// io/sink.cpp -- see sink.h.
#include "sink.h"

#include <mutex>
#include <stdexcept>
#include <system_error>

#include "filesystem.h"
#include "log.h"

namespace paris
{
    sink::sink(const std::string& path, const std::string& prefix, const volume_geometry& vol_geo)
    : path_{path}, vol_geo_(vol_geo)
    {
        try
        {
            if(path_.empty() || path_.back() != '/')
                path_ += '/';
            path_ += prefix;
            if(!create_directory(path))
            {
                log::fatal() << "sink::sink() failed to create output directory at " << path;
                throw stage_construction_error{"sink::sink() failed"};
            }
            handle_ = ddbvf::create(path_, vol_geo_.dim_x, vol_geo_.dim_y, vol_geo_.dim_z);
        }
        catch(const std::system_error& se)
        {
            log::fatal() << "sink::sink(): system error while creating volume: " << se.code() << " - " << se.what();
            throw stage_runtime_error{"sink::sink() failed"};
        }
        catch(const stage_construction_error&)
        {
            throw;
        }
        catch(const std::runtime_error& re)
        {
            log::fatal() << "sink::sink() encountered a runtime error: " << re.what();
            throw stage_construction_error{"sink::sink() failed"};
        }
    }

    auto sink::save(const b200::volume_device_type& v) -> void
    {
        save(v, v.off);
    }

    auto sink::save(const b200::volume_device_type& v, std::uint32_t first_slice) -> void
    {
        try
        {
            // entered concurrently by every device thread: allocation and download run unlocked, only the file
            // is shared (src/sink.cpp:76-81)
            auto host_v = b200::make_volume_host(v.dim_x, v.dim_y, v.dim_z);
            b200::copy_d2h(v, host_v);

            static std::mutex m;
            std::lock_guard<std::mutex> lock{m};
            ddbvf::write(handle_, host_v, first_slice);
        }
        catch(const std::system_error& se)
        {
            log::fatal() << "sink::save(): system error while saving volume: " << se.code() << " - " << se.what();
            throw stage_runtime_error{"sink::save() failed"};
        }
        catch(const std::runtime_error& re)
        {
            log::fatal() << "sink::save(): runtime error while saving volume: " << re.what();
            throw stage_runtime_error{"sink::save() failed"};
        }
    }
}
