// io/his.cpp -- see his.h.
#include "his.h"

#include <cerrno>
#include <cstdio>
#include <cstring>
#include <memory>
#include <system_error>

#include "log.h"

namespace paris
{
    namespace his
    {
        namespace
        {
            constexpr std::size_t file_header_bytes = 68;   // src/his.cpp:44
            constexpr std::uint16_t his_magic = 0x7000;     // src/his.cpp:48

            template <class T>
            auto field(const unsigned char* header, std::size_t offset) -> T
            {
                T v;
                std::memcpy(&v, header + offset, sizeof(T));
                return v;
            }

            auto sample_bytes(std::uint16_t number_type) -> std::size_t
            {
                switch(number_type)   // src/his.cpp:67-75
                {
                    case 2: return 1;     // unsigned char
                    case 4: return 2;     // unsigned short
                    case 32: return 4;    // dword
                    case 64: return 8;    // double
                    case 128: return 4;   // float
                    default: return 0;
                }
            }

            template <class T>
            auto widen(const unsigned char* raw, float* dst, std::size_t n) -> void
            {
                // raw is not necessarily aligned for T
                for(std::size_t i = 0; i < n; ++i)
                {
                    T v;
                    std::memcpy(&v, raw + i * sizeof(T), sizeof(T));
                    dst[i] = static_cast<float>(v);
                }
            }

            struct file_closer { auto operator()(std::FILE* f) const noexcept -> void { if(f) std::fclose(f); } };
        }

        auto read(const std::string& path, file_info& info, const std::function<float*(std::uint32_t)>& frame_buffer)
            -> std::uint32_t
        {
            info = file_info{};
            auto file = std::unique_ptr<std::FILE, file_closer>{std::fopen(path.c_str(), "rb")};
            if(!file)
            {
                log::warning() << "his::load() failed to open file at " << path;
                throw std::system_error{errno, std::generic_category(), path};
            }

            unsigned char header[file_header_bytes] = {};
            if(std::fread(header, 1, file_header_bytes, file.get()) != file_header_bytes
               || field<std::uint16_t>(header, 0) != his_magic)
            {
                log::warning() << "his::load() could not open non-HIS file at " << path;
                return 0;
            }
            if(field<std::uint16_t>(header, 2) != file_header_bytes)
            {
                log::warning() << "his::load() encountered a file header size mismatch at " << path;
                return 0;
            }
            info.image_header_size = field<std::uint16_t>(header, 10);
            const auto ulx = field<std::uint16_t>(header, 12), uly = field<std::uint16_t>(header, 14);
            const auto brx = field<std::uint16_t>(header, 16), bry = field<std::uint16_t>(header, 18);
            info.frames = field<std::uint16_t>(header, 20);
            info.number_type = field<std::uint16_t>(header, 32);
            const auto bytes_per_sample = sample_bytes(info.number_type);
            if(bytes_per_sample == 0 || brx < ulx || bry < uly)
            {
                log::warning() << "his::load() encountered an unsupported data type at " << path;
                return 0;
            }
            info.width = static_cast<std::uint32_t>(brx) - ulx + 1u;    // src/his.cpp:142-147
            info.height = static_cast<std::uint32_t>(bry) - uly + 1u;
            info.valid = true;

            const auto samples = static_cast<std::size_t>(info.width) * info.height;
            auto raw = std::unique_ptr<unsigned char[]>{new unsigned char[samples * bytes_per_sample]};
            auto decoded = 0u;
            for(auto i = 0u; i < info.frames; ++i)
            {
                // every frame is preceded by image_header_size bytes that are skipped (src/his.cpp:150-153)
                if(std::fseek(file.get(), static_cast<long>(info.image_header_size), SEEK_CUR) != 0
                   || std::fread(raw.get(), bytes_per_sample, samples, file.get()) != samples)
                {
                    log::warning() << "his::load() found " << path << " truncated after " << decoded << " of "
                                   << info.frames << " frames";
                    break;
                }
                auto dst = frame_buffer(i);
                if(dst == nullptr)
                    break;
                switch(info.number_type)
                {
                    case 2: widen<std::uint8_t>(raw.get(), dst, samples); break;
                    case 4: widen<std::uint16_t>(raw.get(), dst, samples); break;
                    case 32: widen<std::uint32_t>(raw.get(), dst, samples); break;
                    case 64: widen<double>(raw.get(), dst, samples); break;
                    default: widen<float>(raw.get(), dst, samples); break;
                }
                ++decoded;
            }
            return decoded;
        }

        namespace
        {
            // the checks of read() on the 68-byte file header; info.valid tells whether they passed
            auto parse_header(std::FILE* file, const std::string& path, file_info& info, bool quiet) -> void
            {
                info = file_info{};
                unsigned char header[file_header_bytes] = {};
                if(std::fread(header, 1, file_header_bytes, file) != file_header_bytes
                   || field<std::uint16_t>(header, 0) != his_magic)
                {
                    if(!quiet) log::warning() << "his::load() could not open non-HIS file at " << path;
                    return;
                }
                if(field<std::uint16_t>(header, 2) != file_header_bytes)
                {
                    if(!quiet) log::warning() << "his::load() encountered a file header size mismatch at " << path;
                    return;
                }
                info.image_header_size = field<std::uint16_t>(header, 10);
                const auto ulx = field<std::uint16_t>(header, 12), uly = field<std::uint16_t>(header, 14);
                const auto brx = field<std::uint16_t>(header, 16), bry = field<std::uint16_t>(header, 18);
                info.frames = field<std::uint16_t>(header, 20);
                info.number_type = field<std::uint16_t>(header, 32);
                if(sample_bytes(info.number_type) == 0 || brx < ulx || bry < uly)
                {
                    if(!quiet) log::warning() << "his::load() encountered an unsupported data type at " << path;
                    return;
                }
                info.width = static_cast<std::uint32_t>(brx) - ulx + 1u;
                info.height = static_cast<std::uint32_t>(bry) - uly + 1u;
                info.valid = true;
            }

            auto widen_frame(std::uint16_t number_type, const unsigned char* raw, float* dst, std::size_t samples) -> void
            {
                switch(number_type)
                {
                    case 2: widen<std::uint8_t>(raw, dst, samples); break;
                    case 4: widen<std::uint16_t>(raw, dst, samples); break;
                    case 32: widen<std::uint32_t>(raw, dst, samples); break;
                    case 64: widen<double>(raw, dst, samples); break;
                    default: widen<float>(raw, dst, samples); break;
                }
            }
        }

        auto probe(const std::string& path) -> file_info
        {
            auto info = file_info{};
            auto file = std::unique_ptr<std::FILE, file_closer>{std::fopen(path.c_str(), "rb")};
            if(!file)
            {
                log::warning() << "his::load() failed to open file at " << path;
                throw std::system_error{errno, std::generic_category(), path};
            }
            parse_header(file.get(), path, info, false);
            if(!info.valid)
                return info;
            // complete frames actually present
            std::fseek(file.get(), 0, SEEK_END);
            const auto size = static_cast<unsigned long long>(std::ftell(file.get()));
            const auto frame_bytes = static_cast<unsigned long long>(info.image_header_size)
                                   + static_cast<unsigned long long>(info.width) * info.height * sample_bytes(info.number_type);
            const auto present = size > file_header_bytes ? (size - file_header_bytes) / frame_bytes : 0ull;
            if(present < info.frames)
            {
                log::warning() << "his::load() found " << path << " truncated after " << present << " of " << info.frames
                               << " frames";
                info.frames = static_cast<std::uint32_t>(present);
            }
            return info;
        }

        auto read_frame(const std::string& path, const file_info& info, std::uint32_t frame, float* dst) -> bool
        {
            if(!info.valid || frame >= info.frames || dst == nullptr)
                return false;
            auto file = std::unique_ptr<std::FILE, file_closer>{std::fopen(path.c_str(), "rb")};
            if(!file)
                return false;
            const auto samples = static_cast<std::size_t>(info.width) * info.height;
            const auto bytes = samples * sample_bytes(info.number_type);
            const auto frame_bytes = static_cast<unsigned long long>(info.image_header_size) + bytes;
            const auto pos = file_header_bytes + frame_bytes * frame + info.image_header_size;
            auto raw = std::unique_ptr<unsigned char[]>{new unsigned char[bytes]};
            if(fseeko(file.get(), static_cast<off_t>(pos), SEEK_SET) != 0
               || std::fread(raw.get(), 1, bytes, file.get()) != bytes)
                return false;
            widen_frame(info.number_type, raw.get(), dst, samples);
            return true;
        }

        auto read_frame_u16(const std::string& path, const file_info& info, std::uint32_t frame, std::uint16_t* dst) -> bool
        {
            if(!info.valid || info.number_type != 4 || frame >= info.frames || dst == nullptr)
                return false;
            auto file = std::unique_ptr<std::FILE, file_closer>{std::fopen(path.c_str(), "rb")};
            if(!file)
                return false;
            const auto bytes = static_cast<std::size_t>(info.width) * info.height * sizeof(std::uint16_t);
            const auto frame_bytes = static_cast<unsigned long long>(info.image_header_size) + bytes;
            const auto pos = file_header_bytes + frame_bytes * frame + info.image_header_size;
            // straight from the file into the caller's (pinned) memory: the samples are widened on the GPU
            return fseeko(file.get(), static_cast<off_t>(pos), SEEK_SET) == 0 && std::fread(dst, 1, bytes, file.get()) == bytes;
        }

        auto load(const std::string& path) -> std::vector<image_type>
        {
            auto images = std::vector<image_type>{};
            auto info = file_info{};
            const auto decoded = read(path, info, [&](std::uint32_t) -> float* {
                images.push_back(b200::make_projection_host(info.width, info.height));
                return images.back().buf.get();
            });
            images.resize(decoded);   // (a frame whose samples were cut short is dropped)
            return images;
        }
    }
}
