// io/ddbvf.cpp -- see ddbvf.h.
#include "ddbvf.h"

#include <cerrno>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <system_error>

namespace paris
{
    namespace ddbvf
    {
        namespace
        {
            constexpr std::uint32_t magic = 0xEFDDDAFAu;
            constexpr std::int32_t version = 0x0010;
            constexpr long data_start = 32;

            auto fail(const char* what) -> void
            {
                throw std::system_error{errno != 0 ? errno : EIO, std::generic_category(), what};
            }
        }

        struct handle
        {
            dimensions d{0u, 0u, 0u};
            std::FILE* file = nullptr;
        };

        auto handle_deleter::operator()(handle* h) noexcept -> void
        {
            if(h == nullptr)
                return;
            if(h->file != nullptr)
                std::fclose(h->file);
            delete h;
        }

        auto create(const std::string& path, std::uint32_t dim_x, std::uint32_t dim_y, std::uint32_t dim_z) -> handle_type
        {
            auto h = handle_type{new handle};
            h->d = dimensions{dim_x, dim_y, dim_z};
            const auto full = path + ".ddbvf";
            h->file = std::fopen(full.c_str(), "w+b");
            if(h->file == nullptr)
                fail("ddbvf::create()");

            unsigned char head[data_start] = {};
            const std::uint32_t fields[4] = {dim_x, dim_y, dim_z, static_cast<std::uint32_t>(data_start - 24)};
            std::memcpy(head + 0, &magic, 4);
            std::memcpy(head + 4, &version, 4);
            std::memcpy(head + 8, fields, 16);
            if(std::fwrite(head, 1, sizeof(head), h->file) != sizeof(head) || std::fflush(h->file) != 0)
                fail("ddbvf::create()");
            return h;
        }

        auto open(const std::string& path) -> handle_type
        {
            auto h = handle_type{new handle};
            h->file = std::fopen(path.c_str(), "r+b");
            if(h->file == nullptr)
                fail("ddbvf::open()");
            unsigned char head[24] = {};
            if(std::fread(head, 1, sizeof(head), h->file) != sizeof(head))
                throw std::runtime_error{"Not a ddbvf file: " + path};
            std::uint32_t id = 0;
            std::int32_t ver = 0;
            std::memcpy(&id, head, 4);
            std::memcpy(&ver, head + 4, 4);
            if(id != magic)
                throw std::runtime_error{"Not a ddbvf file: " + path};
            if(ver != version)
                throw std::runtime_error{"Unsupported ddbvf version: " + path};
            std::uint32_t fields[4] = {};
            std::memcpy(fields, head + 8, 16);
            h->d = dimensions{fields[0], fields[1], fields[2]};
            return h;
        }

        auto dims(const handle_type& h) -> dimensions
        {
            return h ? h->d : dimensions{0u, 0u, 0u};
        }

        namespace
        {
            auto seek_slice(handle& h, std::uint32_t first) -> void
            {
                const auto pos = static_cast<std::uint64_t>(h.d.dim_x) * h.d.dim_y * first * sizeof(float);
                if(fseeko(h.file, static_cast<off_t>(data_start) + static_cast<off_t>(pos), SEEK_SET) != 0)
                    fail("ddbvf: seek");
            }
        }

        auto write(handle_type& h, const float* data, std::uint32_t dim_x, std::uint32_t dim_y, std::uint32_t dim_z,
                   std::uint32_t first) -> void
        {
            if(h == nullptr || data == nullptr)
                return;
            if(first >= h->d.dim_z)
                throw std::runtime_error{"ddbvf::write(): Starting position out of bounds"};
            if(dim_x != h->d.dim_x || dim_y != h->d.dim_y || dim_z > h->d.dim_z
               || static_cast<std::uint64_t>(first) + dim_z > h->d.dim_z)
                throw std::runtime_error{"ddbvf::write(): Attempting to save volume to file with wrong dimensions"};
            seek_slice(*h, first);
            const auto n = static_cast<std::size_t>(dim_x) * dim_y * dim_z;
            if(std::fwrite(data, sizeof(float), n, h->file) != n || std::fflush(h->file) != 0)
                fail("ddbvf::write()");
        }

        auto read(handle_type& h, float* data, std::uint32_t first, std::uint32_t count) -> void
        {
            if(h == nullptr || data == nullptr)
                return;
            if(static_cast<std::uint64_t>(first) + count > h->d.dim_z)
                throw std::runtime_error{"ddbvf::read(): slices out of bounds"};
            seek_slice(*h, first);
            const auto n = static_cast<std::size_t>(h->d.dim_x) * h->d.dim_y * count;
            if(std::fread(data, sizeof(float), n, h->file) != n)
                fail("ddbvf::read()");
        }
    }
}
