// b200/backend.cpp -- forwarders from namespace paris::b200 onto the C ABI (include/paris_b200.h).
#include "backend.h"

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>
#include <unordered_map>

namespace paris
{
    namespace b200
    {
        namespace
        {
            [[noreturn]] auto fail(const char* what) -> void
            {
                throw stage_runtime_error{std::string{what} + ": " + paris_b200_last_error()};
            }

            auto must(int rc, const char* what) -> void
            {
                if(rc != PARIS_B200_OK)
                    fail(what);
            }

            // One context per host thread, as the reference keeps one device per thread (src/main.cpp:157-169).
            struct thread_state
            {
                paris_b200_ctx* ctx = nullptr;
                int device = 0;
                ~thread_state()
                {
                    if(ctx != nullptr)
                        paris_b200_ctx_destroy(ctx);
                }
            };

            auto state() -> thread_state&
            {
                thread_local thread_state s;
                return s;
            }

            // per calling thread, like every other piece of backend state: concurrent per-device callers must not
            // plan their slabs with another call's count
            thread_local int g_slab_count = 0;

            // Pinned host buffers are pooled: the reference drops a host projection at the end of every loop
            // iteration (src/main.cpp:100-105) while its upload may still be in flight, and cudaFreeHost would
            // synchronise the device.  A released buffer is reused only after the copy stream has drained -- the copy
            // stream of the thread that released it, so the free lists are PER THREAD (= per device, src/main.cpp:
            // 157-169 runs one host thread per device): a buffer released by one device thread must not be handed to
            // another whose context knows nothing about the upload still reading it.  Ownership is process-wide.
            struct host_pool
            {
                std::mutex m;
                std::unordered_map<float*, std::size_t> owned;          // every pooled pointer -> bytes
                ~host_pool()
                {
                    for(auto& kv : owned)
                        paris_b200_host_free(kv.first);
                }
            };

            auto pool() -> host_pool&
            {
                static host_pool p;
                return p;
            }

            auto free_list() -> std::unordered_multimap<std::size_t, float*>&   // bytes -> pointers this thread released
            {
                thread_local std::unordered_multimap<std::size_t, float*> list;
                return list;
            }

            auto host_acquire(std::size_t bytes, bool zero) -> float*
            {
                auto& mine = free_list();
                auto it = mine.find(bytes);
                if(it != mine.end())
                {
                    auto ptr = it->second;
                    mine.erase(it);
                    // the previous upload (issued by this thread's context) must have left the buffer before it is
                    // overwritten
                    if(state().ctx != nullptr)
                        paris_b200_h2d_wait(state().ctx);
                    if(zero)
                        std::fill_n(ptr, bytes / sizeof(float), 0.f);
                    return ptr;
                }
                void* raw = nullptr;
                must(paris_b200_host_alloc(bytes, zero ? 1 : 0, &raw), "paris_b200_host_alloc");
                auto& p = pool();
                std::lock_guard<std::mutex> lock{p.m};
                p.owned.emplace(static_cast<float*>(raw), bytes);
                return static_cast<float*>(raw);
            }
        }

        auto context() -> paris_b200_ctx*
        {
            auto& s = state();
            if(s.ctx == nullptr)
                must(paris_b200_ctx_create(s.device, &s.ctx), "paris_b200_ctx_create");
            return s.ctx;
        }

        auto set_slab_count(int num) noexcept -> void { g_slab_count = num; }

        auto flush() -> void { must(paris_b200_flush(context()), "paris_b200_flush"); }

        // ---- deleters -------------------------------------------------------------------------------

        auto host_deleter::operator()(float* p) const noexcept -> void
        {
            if(p == nullptr)
                return;
            auto& hp = pool();
            auto bytes = std::size_t{0};
            {
                std::lock_guard<std::mutex> lock{hp.m};
                auto it = hp.owned.find(p);
                if(it == hp.owned.end())
                    return; // borrowed pointers are simply forgotten
                bytes = it->second;
            }
            free_list().emplace(bytes, p);
        }

        auto device_deleter::operator()(float* p) const noexcept -> void
        {
            if(p != nullptr && state().ctx != nullptr)
                paris_b200_dev_free(state().ctx, p);
        }

        auto volume_deleter::operator()(float* p) const noexcept -> void
        {
            if(p != nullptr && state().ctx != nullptr)
                paris_b200_volume_free(state().ctx, p);
        }

        auto filter_deleter::operator()(paris_b200_filter* p) const noexcept -> void
        {
            paris_b200_filter_destroy(p);
        }

        // ---- make_* / copy_* (src/openmp/memory.cpp:33-79, src/cuda/memory.cpp:34-102) -----------------

        auto make_projection_host(std::uint32_t dim_x, std::uint32_t dim_y) -> projection_host_type
        {
            auto ptr = host_acquire(static_cast<std::size_t>(dim_x) * dim_y * sizeof(float), false);
            return projection_host_type{projection_host_buffer_type{ptr}, dim_x, dim_y, 0u, 0.f, metadata{}};
        }

        auto borrow_projection_host(float* pinned, std::uint32_t dim_x, std::uint32_t dim_y) -> projection_host_type
        {
            return projection_host_type{projection_host_buffer_type{pinned}, dim_x, dim_y, 0u, 0.f, metadata{}};
        }

        auto make_projection_device(std::uint32_t dim_x, std::uint32_t dim_y) -> projection_device_type
        {
            void* raw = nullptr;
            must(paris_b200_dev_alloc(context(), static_cast<std::size_t>(dim_x) * dim_y * sizeof(float), &raw),
                 "paris_b200_dev_alloc");
            return projection_device_type{projection_device_buffer_type{static_cast<float*>(raw)}, dim_x, dim_y, 0u, 0.f,
                                          metadata{}};
        }

        auto make_volume_host(std::uint32_t dim_x, std::uint32_t dim_y, std::uint32_t dim_z) -> volume_host_type
        {
            auto ptr = host_acquire(static_cast<std::size_t>(dim_x) * dim_y * dim_z * sizeof(float), true);
            return volume_host_type{volume_host_buffer_type{ptr}, dim_x, dim_y, dim_z, 0u};
        }

        auto make_volume_device(std::uint32_t dim_x, std::uint32_t dim_y, std::uint32_t dim_z) -> volume_device_type
        {
            float* raw = nullptr;
            must(paris_b200_volume_alloc(context(), dim_x, dim_y, dim_z, &raw), "paris_b200_volume_alloc");
            return volume_device_type{volume_device_buffer_type{raw}, dim_x, dim_y, dim_z, 0u};
        }

        auto copy_h2d(const projection_host_type& h_p, projection_device_type& d_p) -> void
        {
            must(paris_b200_proj_h2d(context(), h_p.buf.get(), d_p.buf.get(), h_p.dim_x, h_p.dim_y),
                 "paris_b200_proj_h2d");
            d_p.idx = h_p.idx;
            d_p.phi = h_p.phi;
            d_p.meta = h_p.meta;
        }

        namespace
        {
            // run whatever weight()/apply_filter() recorded on the projection, in place
            auto materialise(const projection_device_type& d_p) -> void
            {
                const auto& m = d_p.meta;
                if(m.weight_pending && m.filter_pending)
                    must(paris_b200_weight_filter(context(), d_p.buf.get(), d_p.dim_x, d_p.dim_y, m.weighting.h_min,
                                                  m.weighting.v_min, m.weighting.d_sd, m.weighting.l_px_row,
                                                  m.weighting.l_px_col, m.filter, m.filter_size),
                         "paris_b200_weight_filter");
                else if(m.weight_pending)
                    must(paris_b200_weight(context(), d_p.buf.get(), d_p.dim_x, d_p.dim_y, m.weighting.h_min,
                                           m.weighting.v_min, m.weighting.d_sd, m.weighting.l_px_row,
                                           m.weighting.l_px_col),
                         "paris_b200_weight");
            }
        }

        auto copy_d2h(const projection_device_type& d_p, projection_host_type& h_p) -> void
        {
            // the pending stages are applied to the device buffer (idempotence is the caller's: the
            // reference never downloads a projection, src/main.cpp:98-105)
            materialise(d_p);
            const_cast<projection_device_type&>(d_p).meta = metadata{};
            must(paris_b200_proj_d2h(context(), d_p.buf.get(), h_p.buf.get(), d_p.dim_x, d_p.dim_y),
                 "paris_b200_proj_d2h");
            h_p.idx = d_p.idx;
            h_p.phi = d_p.phi;
            h_p.meta = metadata{};
        }

        auto copy_h2d(const volume_host_type& h_v, volume_device_type& d_v) -> void
        {
            const auto n = static_cast<std::size_t>(h_v.dim_x) * h_v.dim_y * h_v.dim_z;
            must(paris_b200_vol_h2d(context(), h_v.buf.get(), d_v.buf.get(), n), "paris_b200_vol_h2d");
            d_v.off = h_v.off;
        }

        auto copy_d2h(const volume_device_type& d_v, volume_host_type& h_v) -> void
        {
            const auto n = static_cast<std::size_t>(d_v.dim_x) * d_v.dim_y * d_v.dim_z;
            must(paris_b200_vol_d2h(context(), d_v.buf.get(), h_v.buf.get(), n), "paris_b200_vol_d2h");
            h_v.off = d_v.off;
        }

        // ---- scheduling ---------------------------------------------------------------------------------

        auto make_subvolume_information(const volume_geometry& vol_geo, const detector_geometry& det_geo)
            -> subvolume_info
        {
            const auto vol = paris_b200_volume_geometry{vol_geo.dim_x, vol_geo.dim_y, vol_geo.dim_z,
                                                        vol_geo.l_vx_x, vol_geo.l_vx_y, vol_geo.l_vx_z};
            const auto det = paris_b200_detector_geometry{det_geo.n_row, det_geo.n_col, det_geo.l_px_row,
                                                          det_geo.l_px_col, det_geo.delta_s, det_geo.delta_t,
                                                          det_geo.d_so, det_geo.d_od, det_geo.delta_phi};
            auto out = paris_b200_subvolume_info{};
            if(paris_b200_make_subvolume_information(context(), &vol, &det, g_slab_count, &out) != PARIS_B200_OK)
                throw stage_construction_error{std::string{"make_subvolume_information() failed: "}
                                               + paris_b200_last_error()};
            auto info = subvolume_info{};
            info.geo.dim_x = out.dim_x;
            info.geo.dim_y = out.dim_y;
            info.geo.dim_z = out.dim_z;
            info.geo.remainder = out.remainder;
            info.num = out.num;
            return info;
        }

        // ---- stages ---------------------------------------------------------------------------------------

        auto weight(projection_device_type& p, float h_min, float v_min, float d_sd, float l_px_row, float l_px_col)
            -> void
        {
            if(p.meta.weight_pending || p.meta.filter_pending)
            {
                materialise(p);
                p.meta = metadata{};
            }
            p.meta.weight_pending = true;
            p.meta.weighting = paris_b200_weighting{h_min, v_min, d_sd, l_px_row, l_px_col};
        }

        auto make_filter(std::uint32_t size, float tau) -> filter_buffer_type
        {
            paris_b200_filter* f = nullptr;
            must(paris_b200_filter_create(context(), size, tau, &f), "paris_b200_filter_create");
            return filter_buffer_type{f};
        }

        auto apply_filter(projection_device_type& p, const filter_buffer_type& k, std::uint32_t filter_size,
                          std::uint32_t n_col) -> void
        {
            if(p.meta.filter_pending)
            {
                materialise(p);
                p.meta = metadata{};
            }
            if(p.meta.weight_pending)
            {
                // weight -> filter in a row: keep both pending, the fused kernel does them in one pass
                p.meta.filter_pending = true;
                p.meta.filter = k.get();
                p.meta.filter_size = filter_size;
                return;
            }
            must(paris_b200_apply_filter(context(), p.buf.get(), p.dim_x, p.dim_y, k.get(), filter_size, n_col),
                 "paris_b200_apply_filter");
        }

        auto backproject(const projection_device_type& p, volume_device_type& v, std::uint32_t v_offset,
                         const detector_geometry& det_geo, const volume_geometry& vol_geo,
                         bool enable_roi, const region_of_interest& roi,
                         float sin, float cos, float delta_s, float delta_t) -> void
        {
            const auto det = paris_b200_detector_geometry{det_geo.n_row, det_geo.n_col, det_geo.l_px_row,
                                                          det_geo.l_px_col, det_geo.delta_s, det_geo.delta_t,
                                                          det_geo.d_so, det_geo.d_od, det_geo.delta_phi};
            const auto vol = paris_b200_volume_geometry{vol_geo.dim_x, vol_geo.dim_y, vol_geo.dim_z,
                                                        vol_geo.l_vx_x, vol_geo.l_vx_y, vol_geo.l_vx_z};
            const auto r = paris_b200_roi{roi.x1, roi.x2, roi.y1, roi.y2, roi.z1, roi.z2};

            auto flags = std::uint32_t{0};
            const paris_b200_filter* filter = nullptr;
            const paris_b200_weighting* weighting = nullptr;
            if(p.meta.weight_pending && p.meta.filter_pending)
            {
                flags = PARIS_B200_BP_FUSE_WEIGHT_FILTER;
                filter = p.meta.filter;
                weighting = &p.meta.weighting;
            }
            else if(p.meta.weight_pending)
            {
                materialise(p);
                const_cast<projection_device_type&>(p).meta = metadata{};
            }
            must(paris_b200_backproject(context(), p.buf.get(), p.dim_x, p.dim_y, v.buf.get(), v.dim_x, v.dim_y,
                                        v.dim_z, v_offset, &det, &vol, enable_roi ? 1 : 0, &r, sin, cos, delta_s,
                                        delta_t, flags, filter, weighting),
                 "paris_b200_backproject");
        }

        // ---- devices -----------------------------------------------------------------------------------------

        auto get_devices() -> std::vector<device_handle>
        {
            auto n = 0;
            must(paris_b200_device_count(&n), "paris_b200_device_count");
            auto v = std::vector<device_handle>{};
            for(auto d = 0; d < n; ++d)
                v.push_back(d);
            return v;
        }

        auto set_device(device_handle& device) -> void
        {
            auto& s = state();
            if(s.ctx != nullptr && s.device != device)
            {
                paris_b200_ctx_destroy(s.ctx);
                s.ctx = nullptr;
            }
            s.device = device;
            must(paris_b200_ctx_bind(context()), "paris_b200_ctx_bind");
        }
    }
}
