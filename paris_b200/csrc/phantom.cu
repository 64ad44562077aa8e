// phantom.cu -- synthetic input: analytic cone-beam line integrals of ellipsoids, evaluated in float64
// and rounded once to float32.  Same formulas and conventions as paris_b200/phantom.py (which tests pin
// against this kernel); used by bench.py to build bench-sized raw stacks in milliseconds.
#include "common.cuh"

namespace pb
{
    struct ellipsoid_set
    {
        int n;
        double e[64][8]; // density, a, b, c, x0, y0, z0, theta_deg
    };

    __global__ void __launch_bounds__(256)
    phantom_kernel(float* __restrict__ out, ellipsoid_set set, uint32_t n_row, uint32_t n_col, double l_px_row,
                   double l_px_col, double delta_s, double delta_t, double d_so, double d_sd, float delta_phi,
                   uint32_t first_idx)
    {
        const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
        const uint32_t i = blockIdx.y;
        const uint32_t p = blockIdx.z;
        if(j >= n_row)
            return;
        // the reference's angle: float(idx) * delta_phi in float (src/backprojection.cpp:57)
        const float phi_deg = __fmul_rn(static_cast<float>(first_idx + p), delta_phi);
        const double phi = static_cast<double>(phi_deg) * (3.14159265358979323846 / 180.0);
        const double c = cos(phi), s = sin(phi);
        const double h = -static_cast<double>(n_row) * l_px_row / 2.0 - delta_s * l_px_row + (j + 0.5) * l_px_row;
        const double v = -static_cast<double>(n_col) * l_px_col / 2.0 - delta_t * l_px_col + (i + 0.5) * l_px_col;
        const double sx = -d_so * c, sy = -d_so * s, sz = 0.0;
        const double s_det = d_sd - d_so;
        const double dx = (s_det * c - h * s) - sx;
        const double dy = (s_det * s + h * c) - sy;
        const double dz = v - sz;
        const double norm = sqrt(dx * dx + dy * dy + dz * dz);
        double acc = 0.0;
        for(int e = 0; e < set.n; ++e)
        {
            const double rho = set.e[e][0], a = set.e[e][1], b = set.e[e][2], cz = set.e[e][3];
            const double th = set.e[e][7] * (3.14159265358979323846 / 180.0);
            const double ct = cos(th), st = sin(th);
            const double px = sx - set.e[e][4], py = sy - set.e[e][5], pz = sz - set.e[e][6];
            const double p0 = (px * ct + py * st) / a, p1 = (-px * st + py * ct) / b, p2 = pz / cz;
            const double d0 = (dx * ct + dy * st) / a, d1 = (-dx * st + dy * ct) / b, d2 = dz / cz;
            const double A = d0 * d0 + d1 * d1 + d2 * d2;
            const double B = p0 * d0 + p1 * d1 + p2 * d2;
            const double C = p0 * p0 + p1 * p1 + p2 * p2 - 1.0;
            const double disc = B * B - A * C;
            if(disc > 0.0)
                acc += rho * (2.0 * sqrt(disc) / A) * norm;
        }
        out[(static_cast<size_t>(p) * n_col + i) * n_row + j] = static_cast<float>(acc);
    }

    int launch_phantom(paris_b200_ctx* ctx, const double* h_ellipsoids, uint32_t n, const paris_b200_detector_geometry* det,
                       uint32_t first_idx, uint32_t n_proj, float* d_out)
    {
        ellipsoid_set set{};
        set.n = static_cast<int>(n);
        for(uint32_t e = 0; e < n; ++e)
            for(int c = 0; c < 8; ++c)
                set.e[e][c] = h_ellipsoids[e * 8 + c];
        const double d_sd = std::fabs(static_cast<double>(det->d_so)) + std::fabs(static_cast<double>(det->d_od));
        const dim3 block(256);
        for(uint32_t done = 0; done < n_proj;)
        {
            const uint32_t chunk = std::min<uint32_t>(n_proj - done, 32768u);
            const dim3 grid((det->n_row + 255u) / 256u, det->n_col, chunk);
            phantom_kernel<<<grid, block, 0, ctx->compute>>>(
                d_out + static_cast<size_t>(done) * det->n_col * det->n_row, set, det->n_row, det->n_col,
                det->l_px_row, det->l_px_col, det->delta_s, det->delta_t, det->d_so, d_sd, det->delta_phi,
                first_idx + done);
            PB_CUDA(cudaGetLastError());
            ++ctx->launches;
            done += chunk;
        }
        return PARIS_B200_OK;
    }
}
