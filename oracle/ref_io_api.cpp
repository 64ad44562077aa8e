/*
 * oracle/ref_io_api.cpp -- TEST INFRASTRUCTURE, not product code.
 *
 * C ABI around the reference's UNMODIFIED input/output chain, compiled from where it lies under
 * /root/reference/src (oracle/Makefile):  his.cpp (HIS reader), ddbvf.cpp (volume container), source.cpp
 * (directory walk, angle file, quality filter), sink.cpp, filesystem.cpp, task.cpp -- on the OpenMP backend's
 * host buffers.  program_options.cpp / main.cpp need Boost.Program_options and GLADOS and are not built.
 * Used by tests/test_io.py to pin paris_b200/cpp/io against the reference's own behaviour.
 */
#include <cstdint>
#include <cstring>
#include <exception>
#include <string>

#include "backend.h"
#include "ddbvf.h"
#include "filesystem.h"
#include "his.h"
#include "source.h"
#include "task.h"

namespace
{
    void put(char* dst, std::size_t len, const std::string& text)
    {
        if(dst == nullptr || len == 0)
            return;
        std::strncpy(dst, text.c_str(), len - 1);
        dst[len - 1] = '\0';
    }
}

extern "C"
{
    // src/his.cpp:88 -- decodes up to `capacity` frames into out; dims2 = {width, height}; returns the frame count
    int paris_ref_his_load(const char* path, float* out, std::uint32_t capacity, std::uint32_t* dims2, char* err,
                           std::size_t err_len)
    {
        try
        {
            auto frames = paris::his::load(path);
            dims2[0] = dims2[1] = 0u;
            for(std::size_t i = 0; i < frames.size(); ++i)
            {
                dims2[0] = frames[i].dim_x;
                dims2[1] = frames[i].dim_y;
                if(i < capacity)
                    std::memcpy(out + i * frames[i].dim_x * frames[i].dim_y, frames[i].buf.get(),
                                sizeof(float) * frames[i].dim_x * frames[i].dim_y);
            }
            return static_cast<int>(frames.size());
        }
        catch(const std::exception& e)
        {
            put(err, err_len, e.what());
            return -1;
        }
    }

    // src/ddbvf.cpp:73 + :122 -- create `path`.ddbvf for (dx, dy, dz) and write one volume of vz slices at `first`
    int paris_ref_ddbvf_create_write(const char* path, std::uint32_t dx, std::uint32_t dy, std::uint32_t dz,
                                     const float* data, std::uint32_t vz, std::uint32_t first, char* err,
                                     std::size_t err_len)
    {
        try
        {
            auto h = paris::ddbvf::create(path, dx, dy, dz);
            auto v = paris::backend::make_volume_host(dx, dy, vz);
            std::memcpy(v.buf.get(), data, sizeof(float) * dx * dy * vz);
            paris::ddbvf::write(h, v, first);
            return 0;
        }
        catch(const std::exception& e)
        {
            put(err, err_len, e.what());
            return -1;
        }
    }

    // src/filesystem.cpp:36
    int paris_ref_read_directory(const char* path, char* out, std::size_t out_len, char* err, std::size_t err_len)
    {
        try
        {
            const auto entries = paris::read_directory(path);
            auto joined = std::string{};
            for(const auto& e : entries)
                joined += e + "\n";
            put(out, out_len, joined);
            return static_cast<int>(entries.size());
        }
        catch(const std::exception& e)
        {
            put(err, err_len, e.what());
            return -1;
        }
    }

    // src/source.cpp:74-135 -- drain a source; per projection its idx, phi and first sample.  NB the reference's
    // frame counter is a thread_local static: call this ONCE per loaded copy of the library and thread.
    int paris_ref_source_walk(const char* dir, int enable_angles, const char* angle_file, std::uint32_t quality,
                              std::uint32_t* idx, float* phi, float* first_sample, std::uint32_t capacity, char* err,
                              std::size_t err_len)
    {
        try
        {
            auto src = paris::source{dir, enable_angles != 0, angle_file ? angle_file : "",
                                     static_cast<std::uint16_t>(quality)};
            auto n = 0u;
            while(!src.drained())
            {
                auto p = src.load_next();
                if(n < capacity)
                {
                    idx[n] = p.idx;
                    phi[n] = p.phi;
                    first_sample[n] = p.buf.get()[0];
                }
                ++n;
            }
            return static_cast<int>(n);
        }
        catch(const std::exception& e)
        {
            put(err, err_len, e.what());
            return -1;
        }
    }

    // src/task.cpp:33 -- ids of the tasks make_tasks creates (sanity: one per slab, geometry copied through)
    int paris_ref_make_tasks(std::uint32_t num, std::uint32_t dim_z, std::uint32_t remainder, std::uint32_t* ids,
                             std::uint32_t* dz, std::uint32_t capacity)
    {
        auto po = paris::program_options{};
        auto info = paris::subvolume_info{};
        info.geo = paris::subvolume_geometry{1u, 1u, dim_z, remainder};
        info.num = static_cast<int>(num);
        auto q = paris::make_tasks(po, paris::volume_geometry{}, info);
        auto n = 0u;
        while(!q.empty())
        {
            if(n < capacity)
            {
                ids[n] = q.front().id;
                dz[n] = q.front().subvol_geo.dim_z;
            }
            q.pop();
            ++n;
        }
        return static_cast<int>(n);
    }
}
