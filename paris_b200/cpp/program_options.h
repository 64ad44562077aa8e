// program_options.h -- command line and geometry file of PARIS (/root/reference/src/program_options.h:31-52,
// src/program_options.cpp:36-153), parsed without Boost.Program_options.  Same option names, defaults,
// requirements and messages:
//   --help  --geometry-format  --geometry <file> (required)  --roi  --roi-x1 .. --roi-z2
//   --input <dir>  --output <dir>  --name <prefix = vol>  --angles <file>  --quality <n = 1>
// Geometry file: `key = value` lines (`#` starts a comment) with the nine keys n_row, n_col, l_px_row, l_px_col,
// delta_s, delta_t, d_so, d_od, delta_phi -- all required.
#pragma once

#include <cstdint>
#include <string>

#include "paris_types.h"

namespace paris
{
    struct program_options
    {
        detector_geometry det_geo;

        bool enable_io;
        std::string input_path;
        std::string output_path;
        std::string prefix;

        bool enable_roi;
        region_of_interest roi;

        bool enable_angles;
        std::string angle_path;

        std::uint16_t quality;
    };

    enum class parse_result { ok, exit_success, exit_failure };

    // The parser proper: never exits.  `message` receives what the reference would have printed (help text on
    // exit_success, the error on exit_failure).
    auto parse_program_options(int argc, const char* const* argv, program_options& po, std::string& message) -> parse_result;
    // geometry file only (used by parse_program_options)
    auto parse_geometry_file(const std::string& path, detector_geometry& det_geo, std::string& message) -> bool;

    // the reference's entry point: prints and std::exit()s on --help / --geometry-format / errors
    auto make_program_options(int argc, char** argv) -> program_options;
}
