// io/ddbvf.h -- the DDBVF volume container PARIS writes (/root/reference/src/ddbvf.h:31-47, src/ddbvf.cpp:42-153).
//
// Layout as the reference's create() produces it (src/ddbvf.cpp:73-101):
//   offset 0   u32  0xEFDDDAFA                      (ddbvf_id)
//          4   i32  0x0010                          (ddbvf_version; `constexpr auto` makes it an int, so create()
//                                                    writes FOUR bytes although open() reads a u16 back, :119)
//          8   u32  dim_x   12 u32 dim_y   16 u32 dim_z
//         20   u32  offset  (= 8: distance from the end of this header to the data)
//         24   8 zero bytes
//         32   dim_x*dim_y*dim_z float32, x fastest, z slowest
// Differences from the reference, all on the reading/seeking side: open() reads the 4-byte version
// that create() writes (the reference's 2-byte read leaves every later field misaligned), and slab positions
// are computed in 64 bits (the reference multiplies in uint32_t and wraps beyond 4 GiB, :139-140).
#pragma once

#include <cstdint>
#include <memory>
#include <string>

namespace paris
{
    namespace ddbvf
    {
        struct handle;
        struct handle_deleter { auto operator()(handle* h) noexcept -> void; };
        using handle_type = std::unique_ptr<handle, handle_deleter>;

        struct dimensions { std::uint32_t dim_x, dim_y, dim_z; };

        // creates `path`.ddbvf (the suffix is appended as in src/ddbvf.cpp:75) with its header
        auto create(const std::string& path, std::uint32_t dim_x, std::uint32_t dim_y, std::uint32_t dim_z) -> handle_type;
        // opens an existing file (full path) for reading and writing
        auto open(const std::string& path) -> handle_type;
        auto dims(const handle_type& h) -> dimensions;

        // slices [first, first + dim_z) of the file <- data (dim_x*dim_y*dim_z floats).  Same checks and messages as
        // src/ddbvf.cpp:127-135.
        auto write(handle_type& h, const float* data, std::uint32_t dim_x, std::uint32_t dim_y, std::uint32_t dim_z,
                   std::uint32_t first) -> void;
        auto read(handle_type& h, float* data, std::uint32_t first, std::uint32_t count) -> void;

        // the reference's signature (volume_type = backend::volume_host_type): forwards to the pointer version
        template <class Volume>
        auto write(handle_type& h, const Volume& vol, std::uint32_t first) -> void
        {
            if(h == nullptr || vol.buf == nullptr)
                return;
            write(h, vol.buf.get(), vol.dim_x, vol.dim_y, vol.dim_z, first);
        }
    }
}
