// io_api.cpp -- C entry points over the I/O layer (io/*.h, program_options.h) so that tests/ can drive the HIS
// reader, the DDBVF container, the directory/angle helpers, the option parser and the projection source through
// ctypes.  Errors: -1 with the message in `err` (if non-null).
#include <cstdio>
#include <cstring>
#include <exception>
#include <string>
#include <vector>

#include "io/ddbvf.h"
#include "io/filesystem.h"
#include "io/his.h"
#include "io/source.h"
#include "program_options.h"

namespace
{
    auto put(char* dst, std::size_t len, const std::string& text) -> void
    {
        if(dst == nullptr || len == 0)
            return;
        std::strncpy(dst, text.c_str(), len - 1);
        dst[len - 1] = '\0';
    }

    template <class F>
    auto guarded(char* err, std::size_t err_len, F&& f) -> int
    {
        try
        {
            return f();
        }
        catch(const std::exception& e)
        {
            put(err, err_len, e.what());
            return -1;
        }
    }
}

extern "C"
{
    struct paris_b200_io_options
    {
        paris_b200_detector_geometry det;
        int enable_io, enable_roi, enable_angles;
        paris_b200_roi roi;
        std::uint32_t quality;
        char input_path[512], output_path[512], prefix[128], angle_path[512];
    };

    // geometry of a HIS file without decoding it: {width, height, frames, number_type}; returns 1 if valid
    int paris_b200_io_his_info(const char* path, std::uint32_t* out4, char* err, std::size_t err_len)
    {
        return guarded(err, err_len, [&] {
            auto info = paris::his::file_info{};
            paris::his::read(path, info, [](std::uint32_t) -> float* { return nullptr; });
            out4[0] = info.width;
            out4[1] = info.height;
            out4[2] = info.frames;
            out4[3] = info.number_type;
            return info.valid ? 1 : 0;
        });
    }

    // decodes up to `capacity` frames into out (frame-major); returns the number decoded
    int paris_b200_io_his_read(const char* path, float* out, std::uint32_t capacity, char* err, std::size_t err_len)
    {
        return guarded(err, err_len, [&] {
            auto info = paris::his::file_info{};
            const auto n = paris::his::read(path, info, [&](std::uint32_t i) -> float* {
                return i < capacity ? out + static_cast<std::size_t>(i) * info.width * info.height : nullptr;
            });
            return static_cast<int>(n);
        });
    }

    void* paris_b200_io_ddbvf_create(const char* path, std::uint32_t dx, std::uint32_t dy, std::uint32_t dz, char* err,
                                     std::size_t err_len)
    {
        void* h = nullptr;
        guarded(err, err_len, [&] { h = paris::ddbvf::create(path, dx, dy, dz).release(); return 0; });
        return h;
    }

    void* paris_b200_io_ddbvf_open(const char* path, std::uint32_t* dims3, char* err, std::size_t err_len)
    {
        void* h = nullptr;
        guarded(err, err_len, [&] {
            auto handle = paris::ddbvf::open(path);
            const auto d = paris::ddbvf::dims(handle);
            dims3[0] = d.dim_x;
            dims3[1] = d.dim_y;
            dims3[2] = d.dim_z;
            h = handle.release();
            return 0;
        });
        return h;
    }

    int paris_b200_io_ddbvf_write(void* h, const float* data, std::uint32_t dx, std::uint32_t dy, std::uint32_t dz,
                                  std::uint32_t first, char* err, std::size_t err_len)
    {
        return guarded(err, err_len, [&] {
            auto handle = paris::ddbvf::handle_type{static_cast<paris::ddbvf::handle*>(h)};
            try { paris::ddbvf::write(handle, data, dx, dy, dz, first); }
            catch(...) { (void)handle.release(); throw; }
            (void)handle.release();
            return 0;
        });
    }

    int paris_b200_io_ddbvf_read(void* h, float* data, std::uint32_t first, std::uint32_t count, char* err,
                                 std::size_t err_len)
    {
        return guarded(err, err_len, [&] {
            auto handle = paris::ddbvf::handle_type{static_cast<paris::ddbvf::handle*>(h)};
            try { paris::ddbvf::read(handle, data, first, count); }
            catch(...) { (void)handle.release(); throw; }
            (void)handle.release();
            return 0;
        });
    }

    void paris_b200_io_ddbvf_close(void* h)
    {
        paris::ddbvf::handle_deleter{}(static_cast<paris::ddbvf::handle*>(h));
    }

    // newline-separated canonical paths; returns the number of entries
    int paris_b200_io_read_directory(const char* path, char* out, std::size_t out_len, char* err, std::size_t err_len)
    {
        return guarded(err, err_len, [&] {
            const auto entries = paris::read_directory(path);
            auto joined = std::string{};
            for(const auto& e : entries)
                joined += e + "\n";
            put(out, out_len, joined);
            return static_cast<int>(entries.size());
        });
    }

    int paris_b200_io_create_directory(const char* path, char* err, std::size_t err_len)
    {
        return guarded(err, err_len, [&] { return paris::create_directory(path) ? 1 : 0; });
    }

    int paris_b200_io_read_angles(const char* path, float* out, std::uint32_t capacity, char* err, std::size_t err_len)
    {
        return guarded(err, err_len, [&] {
            const auto a = paris::read_angles(path);
            for(std::size_t i = 0; i < a.size() && i < capacity; ++i)
                out[i] = a[i];
            return static_cast<int>(a.size());
        });
    }

    // 0 = ok, 1 = would exit(EXIT_SUCCESS) (help), 2 = would exit(EXIT_FAILURE); message receives the text
    int paris_b200_io_parse_options(int argc, const char* const* argv, paris_b200_io_options* out, char* message,
                                    std::size_t message_len)
    {
        auto po = paris::program_options{};
        auto text = std::string{};
        const auto r = paris::parse_program_options(argc, argv, po, text);
        put(message, message_len, text);
        if(out != nullptr)
        {
            std::memset(out, 0, sizeof(*out));
            out->det = paris_b200_detector_geometry{po.det_geo.n_row, po.det_geo.n_col, po.det_geo.l_px_row,
                                                    po.det_geo.l_px_col, po.det_geo.delta_s, po.det_geo.delta_t,
                                                    po.det_geo.d_so, po.det_geo.d_od, po.det_geo.delta_phi};
            out->enable_io = po.enable_io;
            out->enable_roi = po.enable_roi;
            out->enable_angles = po.enable_angles;
            out->roi = paris_b200_roi{po.roi.x1, po.roi.x2, po.roi.y1, po.roi.y2, po.roi.z1, po.roi.z2};
            out->quality = po.quality;
            put(out->input_path, sizeof(out->input_path), po.input_path);
            put(out->output_path, sizeof(out->output_path), po.output_path);
            put(out->prefix, sizeof(out->prefix), po.prefix);
            put(out->angle_path, sizeof(out->angle_path), po.angle_path);
        }
        return r == paris::parse_result::ok ? 0 : r == paris::parse_result::exit_success ? 1 : 2;
    }

    // The scan index the group driver plans with (no decoding, no device): per kept projection its frame counter,
    // its angle and whether that came from the angle file; scan3 = {dim_x, dim_y, 1 if every file holds 16-bit
    // samples}.  Returns the count.
    int paris_b200_io_scan_index(const char* dir, int enable_angles, const char* angle_file, std::uint32_t quality,
                                 std::uint32_t* idx, float* phi, int* from_file, std::uint32_t capacity, std::uint32_t* scan3,
                                 char* err, std::size_t err_len)
    {
        return guarded(err, err_len, [&] {
            const auto scan = paris::make_scan_index(dir, enable_angles != 0, angle_file ? angle_file : "",
                                                     static_cast<std::uint16_t>(quality));
            for(auto i = std::size_t{0}; i < scan.frames.size() && i < capacity; ++i)
            {
                idx[i] = scan.frames[i].idx;
                phi[i] = scan.frames[i].phi;
                from_file[i] = scan.frames[i].has_angle ? 1 : 0;
            }
            scan3[0] = scan.dim_x;
            scan3[1] = scan.dim_y;
            scan3[2] = paris::scan_is_u16(scan) ? 1u : 0u;
            return static_cast<int>(scan.frames.size());
        });
    }

    // projection i of that scan: widened to float into out_f32 and / or as 16-bit counts into out_u16 (either may be
    // null); bit 0 / bit 1 of the result say which of the two reads succeeded
    int paris_b200_io_scan_frame(const char* dir, std::uint32_t quality, std::uint32_t i, float* out_f32,
                                 std::uint16_t* out_u16, char* err, std::size_t err_len)
    {
        return guarded(err, err_len, [&] {
            const auto scan = paris::make_scan_index(dir, false, "", static_cast<std::uint16_t>(quality));
            auto got = 0;
            if(out_f32 != nullptr && paris::load_scan_frame(scan, i, out_f32))
                got |= 1;
            if(out_u16 != nullptr && paris::load_scan_frame_u16(scan, i, out_u16))
                got |= 2;
            return got;
        });
    }

    // drains a projection source (pinned host memory: needs a CUDA device) and reports, per projection handed out,
    // its index, angle and whether the angle came from the file; returns the count
    int paris_b200_io_source_walk(const char* dir, int enable_angles, const char* angle_file, std::uint32_t quality,
                                  std::uint32_t* idx, float* phi, int* from_file, float* first_sample,
                                  std::uint32_t capacity, char* err, std::size_t err_len)
    {
        return guarded(err, err_len, [&] {
            auto src = paris::source{dir, enable_angles != 0, angle_file ? angle_file : "",
                                     static_cast<std::uint16_t>(quality)};
            auto n = 0u;
            while(!src.drained())
            {
                auto p = src.load_next();
                if(p.buf == nullptr)
                    break;
                if(n < capacity)
                {
                    idx[n] = p.idx;
                    phi[n] = p.phi;
                    from_file[n] = src.has_angle_for(p.idx) ? 1 : 0;
                    first_sample[n] = p.buf.get()[0];
                }
                ++n;
            }
            return static_cast<int>(n);
        });
    }
}
