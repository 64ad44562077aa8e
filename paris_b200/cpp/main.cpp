// main.cpp -- the `paris_b200` executable: PARIS's driver (/root/reference/src/main.cpp:79-193) on the B200 backend.
//
//   geometry file + command line -> volume geometry (-> ROI) -> z-slab tasks -> one host thread per device, each
//   popping tasks:  source (HIS files) -> load -> weight -> filter -> backproject per projection -> sink (DDBVF).
//
// Same flags and the same output file as the reference.  Two ways through the middle:
//   * one device: the reference's task model (the GLADOS task queue is a mutex-protected std::queue, Boost.Log is
//     io/log.h); the stages run deferred and batched inside the backend (b200/backend.h);
//   * several devices: a reconstruction GROUP (paris_b200_group_*, include/paris_b200.h) instead of independent tasks.
//     The reference lets every device re-read and re-filter the whole scan for each slab it owns (src/main.cpp:93-105);
//     here every device thread reads and filters 1/N of the frames, the filtered detector-row bands travel between the
//     GPUs over peer memory, and every device backprojects all projections into its own slabs.  PARIS_B200_TASKS=1
//     forces the task model, PARIS_B200_GROUP=1 the group (also on one device), PARIS_B200_GROUP_MEMBERS=n runs n
//     members on the devices at hand (members may share a device: testing).
#include <algorithm>
#include <chrono>
#include <condition_variable>
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

namespace
{
    // everybody waits until `count` threads have arrived (reusable)
    class rendezvous
    {
        public:
            explicit rendezvous(std::size_t count) : count_{count} {}
            auto arrive_and_wait() -> void
            {
                auto lock = std::unique_lock<std::mutex>{m_};
                const auto gen = generation_;
                if(++waiting_ == count_)
                {
                    waiting_ = 0;
                    ++generation_;
                    cv_.notify_all();
                }
                else
                    cv_.wait(lock, [&] { return gen != generation_; });
            }

        private:
            std::mutex m_;
            std::condition_variable cv_;
            std::size_t count_, waiting_ = 0, generation_ = 0;
    };

    auto must(int rc, const char* what) -> void
    {
        if(rc != PARIS_B200_OK)
            throw paris::stage_runtime_error{std::string{what} + ": " + paris_b200_last_error()};
    }

    struct pinned_floats
    {
        float* p = nullptr;
        explicit pinned_floats(std::size_t n)
        {
            void* raw = nullptr;
            must(paris_b200_host_alloc(std::max<std::size_t>(n, 1) * sizeof(float), 0, &raw), "paris_b200_host_alloc");
            p = static_cast<float*>(raw);
        }
        pinned_floats(const pinned_floats&) = delete;
        auto operator=(const pinned_floats&) -> pinned_floats& = delete;
        ~pinned_floats() { paris_b200_host_free(p); }
    };

    // One member of the group: reads its share of every round from disk while the device works on the round before,
    // then stores its slabs.  `members` threads run this side by side (src/main.cpp:157-169: one thread per device).
    auto reconstruct_member(std::size_t rank, std::size_t world, int device, const paris::program_options& po,
                            const paris::volume_geometry& vol_geo, const paris::scan_index& scan,
                            const std::vector<float>& angles, std::uint32_t slabs_per_member, paris::sink& sink,
                            std::vector<unsigned char>& handles, rendezvous& meet, std::vector<int>& failed) -> void
    {
        paris_b200_group* group = nullptr;
        auto abandon = [&](const std::exception& e) {
            // a member that cannot go on must still show up at every meeting point, or its peers wait for ever
            failed[rank] = 1;
            paris::log::fatal() << "group member " << rank << ": " << e.what();
        };
        auto cfg = paris_b200_group_config{};
        cfg.rank = static_cast<std::int32_t>(rank);
        cfg.world = static_cast<std::int32_t>(world);
        cfg.det = paris_b200_detector_geometry{po.det_geo.n_row, po.det_geo.n_col, po.det_geo.l_px_row, po.det_geo.l_px_col,
                                               po.det_geo.delta_s, po.det_geo.delta_t, po.det_geo.d_so, po.det_geo.d_od,
                                               po.det_geo.delta_phi};
        cfg.vol_full = paris_b200_volume_geometry{vol_geo.dim_x, vol_geo.dim_y, vol_geo.dim_z, vol_geo.l_vx_x, vol_geo.l_vx_y,
                                                  vol_geo.l_vx_z};
        cfg.enable_roi = po.enable_roi ? 1 : 0;
        cfg.roi = paris_b200_roi{po.roi.x1, po.roi.x2, po.roi.y1, po.roi.y2, po.roi.z1, po.roi.z2};
        cfg.n_proj = static_cast<std::uint32_t>(scan.frames.size());
        cfg.angles_deg = angles.data();
        cfg.slabs_per_member = slabs_per_member;
        cfg.stream_slabs = slabs_per_member > 1u ? 1u : 0u;
        // 16-bit detector files go to the GPU as they are (half the PCIe bytes; the filter kernel widens them with the
        // very conversion src/his.cpp:98-99 does on the host); PARIS_B200_SAMPLES=f32 keeps the host conversion
        const auto* samples_env = std::getenv("PARIS_B200_SAMPLES");
        const auto native_u16 = paris::scan_is_u16(scan) && !(samples_env != nullptr && std::string{samples_env} == "f32");
        const auto sample_bytes = native_u16 ? sizeof(std::uint16_t) : sizeof(float);
        cfg.sample_type = native_u16 ? PARIS_B200_SAMPLES_U16 : PARIS_B200_SAMPLES_F32;
        cfg.exchange = PARIS_B200_EXCHANGE_COPY_ENGINE;

        try
        {
            must(paris_b200_group_create(device, &cfg, &group), "paris_b200_group_create");
            must(paris_b200_group_export(group, handles.data() + rank * PARIS_B200_GROUP_HANDLE_BYTES,
                                         PARIS_B200_GROUP_HANDLE_BYTES), "paris_b200_group_export");
        }
        catch(const std::exception& e) { abandon(e); }
        meet.arrive_and_wait();                                   // every handle is there
        const auto anybody_failed = [&] { return std::any_of(failed.begin(), failed.end(), [](int f) { return f != 0; }); };
        if(!anybody_failed())
        {
            try { must(paris_b200_group_connect(group, handles.data(), PARIS_B200_GROUP_HANDLE_BYTES), "paris_b200_group_connect"); }
            catch(const std::exception& e) { abandon(e); }
        }
        meet.arrive_and_wait();                                   // everybody is connected (or somebody gave up)
        if(!anybody_failed())
        {
            try
            {
                auto info = paris_b200_group_info_t{};
                must(paris_b200_group_info(group, &info), "paris_b200_group_info");
                auto plan = paris_b200_group_plan_t{};
                must(paris_b200_group_plan(&cfg, &plan), "paris_b200_group_plan");
                const auto px = static_cast<std::size_t>(po.det_geo.n_row) * po.det_geo.n_col;
                auto largest = 0u;
                for(auto rd = 0u; rd < plan.rounds; ++rd)
                {
                    auto first = 0u, count = 0u;
                    must(paris_b200_group_share(&plan, static_cast<std::uint32_t>(world), rd, static_cast<std::uint32_t>(rank),
                                                &first, &count), "paris_b200_group_share");
                    largest = std::max(largest, count);
                }
                // two sets of pinned frames: the disk fills one while the upload of the other is in flight
                const auto set_floats = (px * largest * sample_bytes + sizeof(float) - 1u) / sizeof(float);
                pinned_floats frames[2] = {pinned_floats{set_floats}, pinned_floats{set_floats}};
                pinned_floats slabs{static_cast<std::size_t>(info.region_x) * info.region_y * info.z_count};

                must(paris_b200_group_step_open(group, slabs.p), "paris_b200_group_step_open");
                auto loaded = 0u;
                for(auto rd = 0u; rd < plan.rounds; ++rd)
                {
                    auto first = 0u, count = 0u;
                    must(paris_b200_group_share(&plan, static_cast<std::uint32_t>(world), rd, static_cast<std::uint32_t>(rank),
                                                &first, &count), "paris_b200_group_share");
                    auto* set = reinterpret_cast<unsigned char*>(frames[rd & 1u].p);
                    if(rd >= 2u)
                    {
                        auto done = 0;
                        while(done == 0)   // the upload that last used this set
                        {
                            must(paris_b200_group_uploaded(group, rd - 2u, &done), "paris_b200_group_uploaded");
                            if(done == 0)
                                std::this_thread::sleep_for(std::chrono::microseconds{200});
                        }
                    }
                    auto ptrs = std::vector<const float*>(count);
                    for(auto j = 0u; j < count; ++j)
                    {
                        auto* frame = set + px * sample_bytes * j;
                        const auto ok = native_u16 ? paris::load_scan_frame_u16(scan, first + j, reinterpret_cast<std::uint16_t*>(frame))
                                                   : paris::load_scan_frame(scan, first + j, reinterpret_cast<float*>(frame));
                        if(!ok)
                            throw paris::stage_runtime_error{"could not read projection " + std::to_string(first + j)};
                        ptrs[j] = reinterpret_cast<const float*>(frame);
                        ++loaded;
                    }
                    must(paris_b200_group_step_round(group, rd, count > 0u ? ptrs.data() : nullptr, nullptr),
                         "paris_b200_group_step_round");
                }
                must(paris_b200_group_step_finish(group), "paris_b200_group_step_finish");
                must(paris_b200_group_end(group), "paris_b200_group_end");
                paris::log::info() << "device " << device << ": member " << rank + 1u << "/" << world << ", " << loaded
                                   << " of " << scan.frames.size() << (native_u16 ? " 16-bit" : "") << " projections filtered here, slices [" << info.z_first
                                   << ", " << info.z_first + info.z_count << ") in " << info.slabs << (info.slabs == 1u ? " slab" : " slabs")
                                   << ", detector rows [" << info.band_lo << ", " << info.band_hi << ") received";
                sink.save(slabs.p, info.region_x, info.region_y, info.z_count, info.z_first);
            }
            catch(const std::exception& e) { abandon(e); }
        }
        meet.arrive_and_wait();                                   // nobody tears its stack down while a peer may still push into it
        paris_b200_group_destroy(group);
    }

    auto reconstruct_group(const paris::program_options& po, const paris::volume_geometry& vol_geo,
                           const paris::volume_geometry& roi_geo, const std::vector<paris::b200::device_handle>& devices,
                           std::size_t members, std::uint32_t slabs_total, paris::sink& sink) -> void
    {
        const auto scan = paris::make_scan_index(po.input_path, po.enable_angles, po.angle_path, po.quality);
        if(scan.frames.empty())
            throw paris::stage_runtime_error{"no projections found in " + po.input_path};
        if(scan.dim_x != po.det_geo.n_row || scan.dim_y != po.det_geo.n_col)
            throw paris::stage_runtime_error{"the frames in " + po.input_path + " do not have the geometry file's detector size"};
        // the angle of projection i as src/backprojection.cpp:53-57 takes it: from the angle file, else idx * delta_phi
        auto angles = std::vector<float>(scan.frames.size());
        for(auto i = std::size_t{0}; i < scan.frames.size(); ++i)
            angles[i] = scan.frames[i].has_angle ? scan.frames[i].phi : static_cast<float>(scan.frames[i].idx) * po.det_geo.delta_phi;
        members = std::min<std::size_t>(members, roi_geo.dim_z);
        const auto spm = std::max<std::uint32_t>(1u, static_cast<std::uint32_t>((slabs_total + members - 1u) / members));
        paris::log::info() << "Group of " << members << (members == 1 ? " member" : " members") << " on " << devices.size()
                           << (devices.size() == 1 ? " device, " : " devices, ") << scan.frames.size() << " projections, "
                           << spm << (spm == 1u ? " slab" : " slabs") << " per member";

        auto handles = std::vector<unsigned char>(members * PARIS_B200_GROUP_HANDLE_BYTES);
        auto failed = std::vector<int>(members, 0);
        auto meet = rendezvous{members};
        auto threads = std::vector<std::thread>{};
        for(auto r = std::size_t{0}; r < members; ++r)
            threads.emplace_back([&, r] {
                reconstruct_member(r, members, devices[r % devices.size()], po, vol_geo, scan, angles, spm, sink, handles, meet,
                                   failed);
            });
        for(auto& t : threads)
            t.join();
        if(std::any_of(failed.begin(), failed.end(), [](int f) { return f != 0; }))
            throw paris::stage_runtime_error{"the reconstruction group failed"};
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

            auto members = devices.size();
            if(const auto* m = std::getenv("PARIS_B200_GROUP_MEMBERS"))
                members = static_cast<std::size_t>(std::max(1, std::atoi(m)));
            const auto force_tasks = std::getenv("PARIS_B200_TASKS") != nullptr;
            const auto want_group = !force_tasks && (members > 1 || std::getenv("PARIS_B200_GROUP") != nullptr);
            if(want_group)
                reconstruct_group(po, vol_geo, roi_geo, devices, members, static_cast<std::uint32_t>(task_num), sink);
            else if(devices.size() > 1)
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
