// io/sink.h -- the volume sink of /root/reference/src/sink.h:33-46 (src/sink.cpp:36-94): creates the output
// directory and `<path>/<prefix>.ddbvf` for the (ROI-)region and stores every finished z-slab at its offset.
//
// The reference writes every slab at host_v.off, which nothing ever sets; here the slab's first
// slice travels with the device volume (volume::off, set by the task loop) or is given explicitly.
#pragma once

#include <cstdint>
#include <string>

#include "../b200/backend.h"
#include "ddbvf.h"

namespace paris
{
    class sink
    {
        public:
            sink(const std::string& path, const std::string& prefix, const volume_geometry& vol_geo);
            // slab -> host (the pending backprojection batch is flushed behind the download) -> file at v.off
            auto save(const b200::volume_device_type& v) -> void;
            auto save(const b200::volume_device_type& v, std::uint32_t first_slice) -> void;
            // a slab that is already on the host (a group member downloads its slabs itself, behind its kernels)
            auto save(const float* h_slab, std::uint32_t dim_x, std::uint32_t dim_y, std::uint32_t dim_z,
                      std::uint32_t first_slice) -> void;
            auto file_path() const -> std::string { return path_ + ".ddbvf"; }

        private:
            std::string path_;
            ddbvf::handle_type handle_;
            volume_geometry vol_geo_;
    };
}
