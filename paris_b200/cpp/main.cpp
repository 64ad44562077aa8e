// main.cpp -- the `paris_b200` executable: PARIS's driver (/root/reference/src/main.cpp:79-193) on the B200 backend.
//
//   geometry file + command line -> volume geometry (-> ROI) -> z-slab tasks -> one host thread per device, each
//   popping tasks:  source (HIS files) -> load -> weight -> filter -> backproject per projection -> sink (DDBVF).
//
// Same flags, same output file, same task model as the reference.  The GLADOS task queue is a mutex-protected
// std::queue, Boost.Log is io/log.h; the stages run deferred and batched inside the backend (b200/backend.h).
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <exception>
#include <iomanip>
#include <iostream>
#include <mutex>
#include <sstream>
#include <thread>
#include <vector>

#include "io/log.h"
#include "io/sink.h"
#include "io/source.h"
#include "pipeline.h"
#include "program_options.h"
#include "task.h"

namespace
{
    struct task_queue
    {
        std::queue<paris::task> tasks;
        std::mutex m;

        auto pop(paris::task& t) -> bool
        {
            std::lock_guard<std::mutex> lock{m};
            if(tasks.empty())
                return false;
            t = tasks.front();
            tasks.pop();
            return true;
        }
    };

    // src/main.cpp:79-109
    auto reconstruct(task_queue& queue, paris::b200::device_handle device, paris::sink& sink) -> void
    {
        paris::b200::set_device(device);
        auto t = paris::task{};
        while(queue.pop(t))
        {
            const auto last = (t.num - t.id) <= 1u;
            const auto& scan = t.scan;
            auto source = paris::source{scan.input_path, scan.enable_angles, scan.angle_path, scan.quality};

            auto v = paris::make_volume(t.subvol_geo, last);
            const auto offset = t.id * t.subvol_geo.dim_z;
            v.off = offset;

            auto count = 0u;
            while(!source.drained())
            {
                auto p = source.load_next();
                if(p.buf == nullptr)
                    break;
                const auto use_angle = source.has_angle_for(p.idx);
                auto d_p = paris::load(p);
                paris::weight(d_p, scan.det_geo);
                paris::filter(d_p, scan.det_geo);
                paris::backproject(d_p, v, offset, scan.det_geo, t.vol_geo, use_angle, scan.enable_roi, scan.roi);
                ++count;
            }
            paris::log::info() << "device " << device << ": task " << t.id + 1u << "/" << t.num << ", " << count
                               << " projections into slices [" << offset << ", " << offset + v.dim_z << ")";
            sink.save(v);
        }
    }
}

auto main(int argc, char** argv) -> int
{
    std::cout << "PARIS (B200 backend) - " << paris_b200_version() << std::endl;
    auto po = paris::make_program_options(argc, argv);

    try
    {
        const auto vol_geo = paris::calculate_volume_geometry(po.det_geo);
        auto roi_geo = vol_geo;
        if(po.enable_roi)
            roi_geo = paris::apply_roi(vol_geo, po.roi.x1, po.roi.x2, po.roi.y1, po.roi.y2, po.roi.z1, po.roi.z2);
        paris::log::info() << "Volume " << vol_geo.dim_x << " x " << vol_geo.dim_y << " x " << vol_geo.dim_z
                           << " voxels of " << vol_geo.l_vx_x << " mm; region " << roi_geo.dim_x << " x "
                           << roi_geo.dim_y << " x " << roi_geo.dim_z;

        if(po.enable_io)
        {
            const auto start = std::chrono::high_resolution_clock::now();

            auto devices = paris::b200::get_devices();
            if(const auto* limit = std::getenv("PARIS_B200_DEVICES"))
            {
                const auto n = static_cast<std::size_t>(std::max(1, std::atoi(limit)));
                if(n < devices.size())
                    devices.resize(n);
            }
            if(devices.empty())
                throw paris::stage_construction_error{"no usable device"};

            // split the region into z-slabs: as many as the memory of one device demands, and at least one per
            // device (src/cuda/subvolume_information.cpp:78 starts from the device count as well)
            if(const auto* forced = std::getenv("PARIS_B200_SLABS"))   // (testing: force the multi-task path)
                paris::b200::set_slab_count(std::max(1, std::atoi(forced)));
            auto subvol_info = paris::b200::make_subvolume_information(roi_geo, po.det_geo);
            if(static_cast<std::size_t>(subvol_info.num) < devices.size() && roi_geo.dim_z >= devices.size())
            {
                paris::b200::set_slab_count(static_cast<int>(devices.size()));
                subvol_info = paris::b200::make_subvolume_information(roi_geo, po.det_geo);
            }

            auto queue = task_queue{};
            queue.tasks = paris::make_tasks(po, vol_geo, subvol_info);
            const auto task_num = queue.tasks.size();
            paris::log::info() << "Created " << task_num << (task_num == 1 ? " task" : " tasks") << " for "
                               << devices.size() << (devices.size() == 1 ? " device" : " devices");

            auto sink = paris::sink{po.output_path, po.prefix, roi_geo};

            if(devices.size() > 1)
            {
                auto threads = std::vector<std::thread>{};
                auto errors = std::vector<std::exception_ptr>(devices.size());
                for(auto i = std::size_t{0}; i < devices.size(); ++i)
                    threads.emplace_back([&, i] {
                        try { reconstruct(queue, devices[i], sink); }
                        catch(...) { errors[i] = std::current_exception(); }
                    });
                for(auto& t : threads)
                    t.join();
                for(auto& e : errors)
                    if(e)
                        std::rethrow_exception(e);
            }
            else
                reconstruct(queue, devices[0], sink);

            const auto stop = std::chrono::high_resolution_clock::now();
            const auto ms = std::chrono::duration_cast<std::chrono::milliseconds>(stop - start).count();
            auto text = std::ostringstream{};
            text << ms / 60000 << ":" << std::setfill('0') << std::setw(2) << (ms / 1000) % 60 << " minutes ("
                 << ms << " ms)";
            paris::log::info() << "Program terminated. Time elapsed: " << text.str();
            paris::log::info() << "Volume written to " << sink.file_path();
        }
    }
    catch(const paris::stage_construction_error& sce)
    {
        paris::log::fatal() << "main(): Pipeline construction failed: " << sce.what();
        paris::log::fatal() << "Aborting.";
        return EXIT_FAILURE;
    }
    catch(const paris::stage_runtime_error& sre)
    {
        paris::log::fatal() << "main(): Pipeline execution failed: " << sre.what();
        paris::log::fatal() << "Aborting.";
        return EXIT_FAILURE;
    }
    catch(const std::exception& e)
    {
        paris::log::fatal() << "main(): " << e.what();
        return EXIT_FAILURE;
    }
    return EXIT_SUCCESS;
}
