// io/his.h -- reader for the HIS detector format (Perkin-Elmer/Varex frame files), the projection input of
// PARIS (/root/reference/src/his.h:31-38, src/his.cpp:42-198).
//
// File layout as the reference reads it (src/his.cpp:49-66, :113-126; all little endian, no padding):
//   offset  0  u16 file_type (= 0x7000)       2  u16 header_size (= 68)     4  u16 header_version
//           6  u32 file_size                  10  u16 image_header_size
//          12  u16 ulx   14  u16 uly   16  u16 brx   18  u16 bry      (inclusive bounding rectangle)
//          20  u16 frame_number               22  u16 correction
//          24  f64 integration_time           32  u16 number_type           34  34 reserved bytes
//   then, per frame: image_header_size bytes to skip (:152-153), then width*height samples of number_type
//   (2 = u8, 4 = u16, 32 = u32, 64 = f64, 128 = f32; :67-75), row-major, converted to float (:98-99).
#pragma once

#include <cstdint>
#include <functional>
#include <string>
#include <vector>

#include "../b200/backend.h"

namespace paris
{
    namespace his
    {
        using image_type = b200::projection_host_type;

        // Same contract as the reference's his::load: every frame of the file as a host projection (pinned
        // memory here, so the upload that follows is asynchronous).  Throws std::system_error if the file
        // cannot be opened; returns an empty vector for a file that is not a valid/supported HIS file.
        auto load(const std::string& path) -> std::vector<image_type>;

        struct file_info
        {
            std::uint32_t width = 0, height = 0, frames = 0;
            std::uint16_t number_type = 0;
            std::uint16_t image_header_size = 0;
            bool valid = false;
        };

        // The decoding core without any allocation policy: `frame_buffer(i)` must return room for
        // width*height floats for frame i (or nullptr to stop).  Returns the number of frames decoded.
        // Used by load() and, with plain memory, by the CPU-side format tests.
        auto read(const std::string& path, file_info& info, const std::function<float*(std::uint32_t)>& frame_buffer)
            -> std::uint32_t;

        // Header only: what read() would find (info.frames = the COMPLETE frames the file really holds; a truncated
        // file counts fewer than its header says).  info.valid is false for anything read() would reject.
        auto probe(const std::string& path) -> file_info;
        // Frame `frame` of a file probe() accepted, widened to float into dst (width*height floats).  False if the
        // frame cannot be read.  Lets several readers share one scan without decoding each other's frames.
        auto read_frame(const std::string& path, const file_info& info, std::uint32_t frame, float* dst) -> bool;
        // The same for a file of 16-bit samples (number_type 4), WITHOUT widening: width*height uint16 values, as the
        // detector wrote them (little endian, like the host).  False for any other sample type.
        auto read_frame_u16(const std::string& path, const file_info& info, std::uint32_t frame, std::uint16_t* dst) -> bool;
    }
}
