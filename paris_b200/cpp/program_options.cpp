// program_options.cpp -- see program_options.h.
#include "program_options.h"

#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <set>
#include <sstream>

namespace paris
{
    namespace
    {
        const char* const help_text =
            "General options:\n"
            "  --help                 produce a help message\n"
            "  --geometry-format      Display geometry file format\n"
            "\nGeometry options:\n"
            "  --geometry arg         Path to geometry file\n"
            "  --roi                  Region of interest switch (optional)\n"
            "\nInput/output options:\n"
            "  --input arg            Path to projections (optional)\n"
            "  --output arg           Output directory for the reconstructed volume (optional)\n"
            "  --name arg (=vol)      Name of the reconstructed volume (optional)\n"
            "\nReconstruction options:\n"
            "  --angles arg           Path to projection angles (optional)\n"
            "  --quality arg (=1)     Quality setting (optional)\n"
            "\nRegion of Interest options:\n"
            "  --roi-x1 arg           leftmost coordinate\n"
            "  --roi-x2 arg           rightmost coordinate\n"
            "  --roi-y1 arg           uppermost coordinate\n"
            "  --roi-y2 arg           lowest coordinate\n"
            "  --roi-z1 arg           uppermost slice\n"
            "  --roi-z2 arg           lowest slice\n";

        const char* const geometry_text =
            "Geometry file:\n"
            "  --n_row arg            [integer] number of pixels per detector row (= projection width)\n"
            "  --n_col arg            [integer] number of pixels per detector column (= projection height)\n"
            "  --l_px_row arg         [float] horizontal pixel size (= distance between pixel centers) in mm\n"
            "  --l_px_col arg         [float] vertical pixel size (= distance between pixel centers) in mm\n"
            "  --delta_s arg          [float] horizontal detector offset in pixels\n"
            "  --delta_t arg          [float] vertical detector offset in pixels\n"
            "  --d_so arg             [float] distance between object (= center of rotation) and source in mm\n"
            "  --d_od arg             [float] distance between object (= center of rotation) and detector in mm\n"
            "  --delta_phi arg        [float] angle step between two successive projections in degrees\n";

        auto trim(const std::string& s) -> std::string
        {
            const auto b = s.find_first_not_of(" \t\r\n");
            if(b == std::string::npos)
                return std::string{};
            return s.substr(b, s.find_last_not_of(" \t\r\n") - b + 1);
        }

        auto to_u32(const std::string& name, const std::string& text, std::uint32_t& out, std::string& message) -> bool
        {
            errno = 0;
            char* end = nullptr;
            const auto v = std::strtoull(text.c_str(), &end, 10);
            if(text.empty() || text[0] == '-' || end == text.c_str() || *end != '\0' || errno != 0 || v > 0xffffffffull)
            {
                message = "the argument ('" + text + "') for option '--" + name + "' is invalid";
                return false;
            }
            out = static_cast<std::uint32_t>(v);
            return true;
        }

        auto to_float(const std::string& name, const std::string& text, float& out, std::string& message) -> bool
        {
            errno = 0;
            char* end = nullptr;
            const auto v = std::strtof(text.c_str(), &end);
            if(text.empty() || end == text.c_str() || *end != '\0')
            {
                message = "the argument ('" + text + "') for option '--" + name + "' is invalid";
                return false;
            }
            out = v;
            return true;
        }

        auto missing(const std::string& name) -> std::string
        {
            return "the option '--" + name + "' is required but missing";
        }
    }

    auto parse_geometry_file(const std::string& path, detector_geometry& det_geo, std::string& message) -> bool
    {
        auto values = std::map<std::string, std::string>{};
        auto file = std::ifstream{path.c_str()};
        if(file)
        {
            auto line = std::string{};
            while(std::getline(file, line))
            {
                line = trim(line.substr(0, line.find('#')));
                if(line.empty())
                    continue;
                const auto eq = line.find('=');
                if(eq == std::string::npos)
                {
                    message = "the options configuration file contains an invalid line '" + line + "'";
                    return false;
                }
                const auto key = trim(line.substr(0, eq));
                static const std::set<std::string> known = {"n_row", "n_col", "l_px_row", "l_px_col", "delta_s",
                                                            "delta_t", "d_so", "d_od", "delta_phi"};
                if(known.count(key) == 0)
                {
                    message = "unrecognised option '" + key + "'";
                    return false;
                }
                values[key] = trim(line.substr(eq + 1));
            }
        }
        // (an unreadable file leaves every key missing, as in src/program_options.cpp:139-142)
        struct { const char* key; std::uint32_t* u; float* f; } const fields[] = {
            {"n_row", &det_geo.n_row, nullptr},      {"n_col", &det_geo.n_col, nullptr},
            {"l_px_row", nullptr, &det_geo.l_px_row}, {"l_px_col", nullptr, &det_geo.l_px_col},
            {"delta_s", nullptr, &det_geo.delta_s},   {"delta_t", nullptr, &det_geo.delta_t},
            {"d_so", nullptr, &det_geo.d_so},         {"d_od", nullptr, &det_geo.d_od},
            {"delta_phi", nullptr, &det_geo.delta_phi}};
        for(const auto& f : fields)
        {
            const auto it = values.find(f.key);
            if(it == values.end())
            {
                message = missing(f.key);
                return false;
            }
            if(f.u != nullptr ? !to_u32(f.key, it->second, *f.u, message) : !to_float(f.key, it->second, *f.f, message))
                return false;
        }
        return true;
    }

    auto parse_program_options(int argc, const char* const* argv, program_options& po, std::string& message) -> parse_result
    {
        po = program_options{};
        po.prefix = "vol";
        po.quality = 1;
        message.clear();

        // option -> takes a value
        static const std::map<std::string, bool> known = {
            {"help", false}, {"geometry-format", false}, {"geometry", true}, {"roi", false},
            {"roi-x1", true}, {"roi-x2", true}, {"roi-y1", true}, {"roi-y2", true}, {"roi-z1", true}, {"roi-z2", true},
            {"input", true}, {"output", true}, {"name", true}, {"angles", true}, {"quality", true}};
        auto given = std::map<std::string, std::string>{};
        for(auto i = 1; i < argc; ++i)
        {
            const auto arg = std::string{argv[i]};
            if(arg.size() < 3 || arg.compare(0, 2, "--") != 0)
            {
                message = "too many positional options have been specified on the command line";
                return parse_result::exit_failure;
            }
            auto name = arg.substr(2);
            auto value = std::string{};
            auto has_value = false;
            const auto eq = name.find('=');
            if(eq != std::string::npos)
            {
                value = name.substr(eq + 1);
                name = name.substr(0, eq);
                has_value = true;
            }
            const auto it = known.find(name);
            if(it == known.end())
            {
                message = "unrecognised option '--" + name + "'";
                return parse_result::exit_failure;
            }
            if(it->second && !has_value)
            {
                if(i + 1 >= argc)
                {
                    message = "the required argument for option '--" + name + "' is missing";
                    return parse_result::exit_failure;
                }
                value = argv[++i];
            }
            else if(!it->second && has_value)
            {
                message = "option '--" + name + "' does not take any arguments";
                return parse_result::exit_failure;
            }
            given[name] = value;
        }

        if(given.count("help"))
        {
            message = help_text;
            return parse_result::exit_success;
        }
        if(given.count("geometry-format"))
        {
            message = geometry_text;
            return parse_result::exit_success;
        }

        const auto has = [&](const char* n) { return given.count(n) != 0; };
        if(has("input") || has("output"))
        {
            po.enable_io = true;
            for(const auto* n : {"input", "output"})
                if(!has(n))
                {
                    message = missing(n);
                    return parse_result::exit_failure;
                }
            po.input_path = given["input"];
            po.output_path = given["output"];
        }
        if(has("roi"))
        {
            po.enable_roi = true;
            for(const auto* n : {"roi-x1", "roi-x2", "roi-y1", "roi-y2", "roi-z1", "roi-z2"})
                if(!has(n))
                {
                    message = missing(n);
                    return parse_result::exit_failure;
                }
        }
        struct { const char* name; std::uint32_t* dst; } const roi_fields[] = {
            {"roi-x1", &po.roi.x1}, {"roi-x2", &po.roi.x2}, {"roi-y1", &po.roi.y1},
            {"roi-y2", &po.roi.y2}, {"roi-z1", &po.roi.z1}, {"roi-z2", &po.roi.z2}};
        for(const auto& f : roi_fields)
            if(has(f.name) && !to_u32(f.name, given[f.name], *f.dst, message))
                return parse_result::exit_failure;
        if(has("angles"))
        {
            po.enable_angles = true;
            po.angle_path = given["angles"];
        }
        if(has("name"))
            po.prefix = given["name"];
        if(has("quality"))
        {
            auto q = std::uint32_t{};
            if(!to_u32("quality", given["quality"], q, message) || q > 0xffffu)
            {
                message = "the argument ('" + given["quality"] + "') for option '--quality' is invalid";
                return parse_result::exit_failure;
            }
            po.quality = static_cast<std::uint16_t>(q);
        }
        if(!has("geometry"))
        {
            message = missing("geometry");
            return parse_result::exit_failure;
        }
        if(!parse_geometry_file(given["geometry"], po.det_geo, message))
            return parse_result::exit_failure;
        return parse_result::ok;
    }

    auto make_program_options(int argc, char** argv) -> program_options
    {
        auto po = program_options{};
        auto message = std::string{};
        switch(parse_program_options(argc, argv, po, message))
        {
            case parse_result::exit_success:
                std::cout << message << std::endl;
                std::exit(EXIT_SUCCESS);
            case parse_result::exit_failure:
                std::cerr << message << std::endl;
                std::exit(EXIT_FAILURE);
            default:
                break;
        }
        return po;
    }
}
