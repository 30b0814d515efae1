// ek_io.cu -- diagnostics and field dumps in the reference's formats.
//
// Replaces current() (LBM.cu:2674-2710), the reduction of record_umax()
// (LBM.cu:2712-2753), save_data_tecplot() (LBM.cu:2492-2565),
// save_data_end() (LBM.cu:2567-2627) and read_data() (LBM.cu:2629-2671); adds an
// exact binary checkpoint of the full state (the text restart keeps 6 decimals).  The diagnostics are device reductions
// (the reference copies three full fields to the host for each); the dumps
// keep the reference's text formats, including the dump-time linear
// extrapolation of rho, c+, c-, u onto the wall planes (LBM.cu:2527-2542).
#include <stdio.h>
#include <string.h>

#include <vector>

#include "ek_handle.h"

namespace {

// per-row partial sums of (c+ - c-)*Ez on the top plane, charges extrapolated
__global__ void k_wall_current(EkConst c, const double *ch, const double *chn, const double *ez, double *partial)
{
    const int y = blockIdx.x;
    __shared__ double sh[128];
    double acc = 0.0;
    for (int x = threadIdx.x; x < c.NX; x += blockDim.x) {
        const size_t col = (size_t)y * c.PX + x;
        const size_t i1 = (size_t)(c.NZ - 1) * c.plane + col, i2 = (size_t)(c.NZ - 2) * c.plane + col,
                     i3 = (size_t)(c.NZ - 3) * c.plane + col;
        const double ct = 2.0 * ch[i2] - ch[i3];
        const double cnt = 2.0 * chn[i2] - chn[i3];
        acc += (ct - cnt) * ez[i1];
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[y] = sh[0];
}

__global__ void k_max_uz(EkConst c, const double *uz, double *partial)
{
    const int y = blockIdx.x, z = blockIdx.y;
    __shared__ double sh[128];
    double m = 0.0;
    for (int x = threadIdx.x; x < c.NX; x += blockDim.x) m = fmax(m, uz[(size_t)z * c.plane + (size_t)y * c.PX + x]);
    sh[threadIdx.x] = m;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[(size_t)z * c.NY + y] = sh[0];
}

ek_status fetch_all(ek_handle *h, EkHostFields &H)
{
    const size_t cells = (size_t)h->c.NX * h->c.NY * h->c.NZ;
    for (int k = 0; k < EK_NFIELDS; ++k) {
        H.f[k].resize(cells);
        ek_status st = ek_get_field(h, k, H.f[k].data(), 0);
        if (st != EK_OK) return st;
    }
    ek_io_extrapolate_walls(H, h->c.NX, h->c.NY, h->c.NZ);
    return EK_OK;
}

EkDumpGrid grid_of(ek_handle *h)
{
    EkDumpGrid g;
    g.NX = h->c.NX; g.NY = h->c.NY; g.NZ = h->c.NZ;
    g.dx = h->p.dx; g.dy = h->p.dy; g.dz = h->p.dz;
    return g;
}

}  // namespace

// dump-time linear extrapolation of rho, c+, c-, u onto the wall planes (LBM.cu:2527-2542)
void ek_io_extrapolate_walls(EkHostFields &H, int NX, int NY, int NZ)
{
    auto idx = [&](int x, int y, int z) { return (size_t)NX * ((size_t)NY * z + y) + x; };
    const int ext[6] = {EK_RHO, EK_CHARGE, EK_CHARGEN, EK_UX, EK_UY, EK_UZ};
    for (int y = 0; y < NY; ++y)
        for (int x = 0; x < NX; ++x)
            for (int k = 0; k < 6; ++k) {
                std::vector<double> &a = H.f[ext[k]];
                a[idx(x, y, 0)] = 2.0 * a[idx(x, y, 1)] - a[idx(x, y, 2)];
                a[idx(x, y, NZ - 1)] = 2.0 * a[idx(x, y, NZ - 2)] - a[idx(x, y, NZ - 3)];
            }
}

// save_data_tecplot (LBM.cu:2492-2565) from host arrays of the whole domain
bool ek_io_write_tecplot(const char *path, const EkDumpGrid &g, const EkHostFields &H, double time, int append, int first)
{
    FILE *f = fopen(path, append ? "ab" : "wb");
    if (!f) return false;
    const int NX = g.NX, NY = g.NY, NZ = g.NZ;
    if (first)
        fprintf(f, "%s\n",
                "VARIABLES=\"x\",\"y\",\"z\",\"u\",\"v\",\"w\",\"p\",\"charge\",\"neg charge\",\"phi\",\"Ex\",\"Ey\",\"Ez\",\"Temperature\"");
    fprintf(f, "\n");
    fprintf(f, "ZONE T=\"t=%g\", F=POINT, I = %d, J = %d, K = %d\n", time, NX, NY, NZ);
    for (int z = 0; z < NZ; ++z)
        for (int y = 0; y < NY; ++y)
            for (int x = 0; x < NX; ++x) {
                const size_t i = (size_t)NX * ((size_t)NY * z + y) + x;
                fprintf(f, "%g %g %g %g %g %g %g %g %10.6f %10.6f %10.6f %10.6f %10.6f %10.6f\n", g.dx * x, g.dy * y,
                        g.dz * z, H.f[EK_UX][i], H.f[EK_UY][i], H.f[EK_UZ][i], H.f[EK_RHO][i], H.f[EK_CHARGE][i],
                        H.f[EK_CHARGEN][i], H.f[EK_PHI][i], H.f[EK_EX][i], H.f[EK_EY][i], H.f[EK_EZ][i], H.f[EK_T][i]);
            }
    fclose(f);
    return true;
}

// save_data_end (LBM.cu:2567-2627)
bool ek_io_write_end(const char *path, const EkDumpGrid &g, const EkHostFields &H, double time)
{
    FILE *f = fopen(path, "wb");
    if (!f) return false;
    const size_t cells = (size_t)g.NX * g.NY * g.NZ;
    for (size_t i = 0; i < cells; ++i)
        fprintf(f, "%10.6f %10.6f %10.6f %10.6f %10.6f %10.6f %10.6f %10.6f %10.6f %10.6f %10.6f %10.6f\n", time,
                H.f[EK_UX][i], H.f[EK_UY][i], H.f[EK_UZ][i], H.f[EK_RHO][i], H.f[EK_CHARGE][i], H.f[EK_CHARGEN][i],
                H.f[EK_PHI][i], H.f[EK_EX][i], H.f[EK_EY][i], H.f[EK_EZ][i], H.f[EK_T][i]);
    fclose(f);
    return true;
}

extern "C" {

ek_status ek_wall_current(ek_handle *h, double *current)
{
    if (!h || !current) return EK_ERR_INVALID;
    if (!h->allocated) return EK_ERR_STATE;
    DeviceGuard g(h->device);
    const EkConst &c = h->c;
    if (h->efield_stale) {
        ek_launch_efield(c, h->fld[EK_PHI], h->fld[EK_EX], h->fld[EK_EY], h->fld[EK_EZ], h->stream);
        h->efield_stale = false;
    }
    double *partial = nullptr;
    EK_CUDA(h, cudaMalloc((void **)&partial, c.NY * sizeof(double)));
    k_wall_current<<<c.NY, 128, 0, h->stream>>>(c, h->fld[EK_CHARGE], h->fld[EK_CHARGEN], h->fld[EK_EZ], partial);
    std::vector<double> host(c.NY);
    cudaError_t e = cudaMemcpyAsync(host.data(), partial, c.NY * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(partial);
    EK_CUDA(h, e);
    double I = 0;
    for (int y = 0; y < c.NY; ++y) I += host[y];
    *current = I * h->p.K * h->p.dz * h->p.dz;  // LBM.cu:2708
    return EK_OK;
}

ek_status ek_max_uz(ek_handle *h, double *umax)
{
    if (!h || !umax) return EK_ERR_INVALID;
    if (!h->allocated) return EK_ERR_STATE;
    DeviceGuard g(h->device);
    const EkConst &c = h->c;
    const size_t n = (size_t)c.NY * c.NZ;
    double *partial = nullptr;
    EK_CUDA(h, cudaMalloc((void **)&partial, n * sizeof(double)));
    k_max_uz<<<dim3(c.NY, c.NZ), 128, 0, h->stream>>>(c, h->fld[EK_UZ], partial);
    std::vector<double> host(n);
    cudaError_t e = cudaMemcpyAsync(host.data(), partial, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(partial);
    EK_CUDA(h, e);
    double m = 0;  // LBM.cu:2717: starts from 0
    for (size_t i = 0; i < n; ++i) m = host[i] > m ? host[i] : m;
    *umax = m;
    return EK_OK;
}

ek_status ek_save_data_tecplot(ek_handle *h, const char *path, double time, int append, int first)
{
    if (!h || !path) return EK_ERR_INVALID;
    EkHostFields H;
    ek_status st = fetch_all(h, H);
    if (st != EK_OK) return st;
    if (!ek_io_write_tecplot(path, grid_of(h), H, time, append, first)) {
        ek_set_error(h, std::string("cannot open ") + path);
        return EK_ERR_INVALID;
    }
    return EK_OK;
}

ek_status ek_save_data_end(ek_handle *h, const char *path, double time)
{
    if (!h || !path) return EK_ERR_INVALID;
    EkHostFields H;
    ek_status st = fetch_all(h, H);
    if (st != EK_OK) return st;
    if (!ek_io_write_end(path, grid_of(h), H, time)) {
        ek_set_error(h, std::string("cannot open ") + path);
        return EK_ERR_INVALID;
    }
    return EK_OK;
}

// read_data() (LBM.cu:2629-2671): the text file written by save_data_end(), twelve
// columns per cell in scalar_index order (time ux uy uz rho c+ c- phi Ex Ey Ez T).  As in
// the reference the macroscopic arrays are restored and the caller re-creates the
// populations with ek_init_equilibrium() (main.cu:161-176); E is taken from the file.
}  // extern "C"

// parser of the save_data_end text format; false with `err` set on failure
bool ek_io_read_end(const char *path, size_t cells, EkHostFields &H, double *time, std::string &err)
{
    FILE *f = fopen(path, "r");
    if (!f) { err = std::string("cannot open ") + path; return false; }
    for (int k = 0; k < EK_NFIELDS; ++k) H.f[k].resize(cells);
    double t = 0.0;
    for (size_t i = 0; i < cells; ++i) {
        const int n = fscanf(f, "%lf %lf %lf %lf %lf %lf %lf %lf %lf %lf %lf %lf", &t, &H.f[EK_UX][i], &H.f[EK_UY][i],
                             &H.f[EK_UZ][i], &H.f[EK_RHO][i], &H.f[EK_CHARGE][i], &H.f[EK_CHARGEN][i], &H.f[EK_PHI][i],
                             &H.f[EK_EX][i], &H.f[EK_EY][i], &H.f[EK_EZ][i], &H.f[EK_T][i]);
        if (n != 12) {
            fclose(f);
            err = std::string(path) + ": truncated or malformed restart file (cell " + std::to_string(i) + ")";
            return false;
        }
    }
    fclose(f);
    if (time) *time = t;
    return true;
}

extern "C" {

ek_status ek_read_data(ek_handle *h, const char *path, double *time)
{
    if (!h || !path) return EK_ERR_INVALID;
    EkHostFields H;
    std::string err;
    if (!ek_io_read_end(path, (size_t)h->c.NX * h->c.NY * h->c.NZ, H, time, err)) {
        ek_set_error(h, err);
        return EK_ERR_INVALID;
    }
    const double *ptr[EK_NFIELDS];
    for (int k = 0; k < EK_NFIELDS; ++k) ptr[k] = H.f[k].data();
    return ek_set_fields(h, ptr, 0);
}

// ---- exact binary checkpoint: header, 11 fields, c+ - c-, the four population sets in the
// reference's natural order (27 x cells, pre-collision values) -- independent of the
// in-place layout, the A-A parity and the slab/ghost padding of the run that wrote it.

ek_status ek_checkpoint_save(ek_handle *h, const char *path, double time)
{
    if (!h || !path) return EK_ERR_INVALID;
    if (!h->pops_ready) { ek_set_error(h, "ek_checkpoint_save before the populations exist"); return EK_ERR_STATE; }
    FILE *f = fopen(path, "wb");
    if (!f) { ek_set_error(h, std::string("cannot open ") + path); return EK_ERR_INVALID; }
    EkCkptHeader hd;
    memcpy(hd.magic, "EKB200C1", 8);
    hd.NX = h->c.NX; hd.NY = h->c.NY; hd.NZ = h->c.NZ; hd.nfields = EK_NFIELDS;
    hd.steps = h->steps; hd.time = time;
    bool ok = fwrite(&hd, sizeof(hd), 1, f) == 1;
    const size_t cells = (size_t)h->c.NX * h->c.NY * h->c.NZ;
    std::vector<double> buf(cells * 27);
    ek_status st = EK_OK;
    for (int k = 0; k < EK_NFIELDS && ok && st == EK_OK; ++k) {
        st = ek_get_field(h, k, buf.data(), 0);
        ok = st == EK_OK && fwrite(buf.data(), sizeof(double), cells, f) == cells;
    }
    for (int s = 0; s < 4 && ok && st == EK_OK; ++s) {
        st = ek_get_populations(h, s, buf.data(), 0);
        ok = st == EK_OK && fwrite(buf.data(), sizeof(double), cells * 27, f) == cells * 27;
    }
    fclose(f);
    if (st != EK_OK) return st;
    if (!ok) { ek_set_error(h, std::string("short write to ") + path); return EK_ERR_INVALID; }
    return EK_OK;
}

ek_status ek_checkpoint_load(ek_handle *h, const char *path, double *time)
{
    if (!h || !path) return EK_ERR_INVALID;
    FILE *f = fopen(path, "rb");
    if (!f) { ek_set_error(h, std::string("cannot open ") + path); return EK_ERR_INVALID; }
    EkCkptHeader hd;
    if (fread(&hd, sizeof(hd), 1, f) != 1 || memcmp(hd.magic, "EKB200C1", 8) != 0 || hd.NX != h->c.NX ||
        hd.NY != h->c.NY || hd.NZ != h->c.NZ || hd.nfields != EK_NFIELDS) {
        fclose(f);
        ek_set_error(h, std::string(path) + ": not a checkpoint of this grid");
        return EK_ERR_INVALID;
    }
    const size_t cells = (size_t)h->c.NX * h->c.NY * h->c.NZ;
    EkHostFields H;
    bool ok = true;
    for (int k = 0; k < EK_NFIELDS && ok; ++k) {
        H.f[k].resize(cells);
        ok = fread(H.f[k].data(), sizeof(double), cells, f) == cells;
    }
    ek_status st = EK_OK;
    if (ok) {
        const double *ptr[EK_NFIELDS];
        for (int k = 0; k < EK_NFIELDS; ++k) ptr[k] = H.f[k].data();
        st = ek_set_fields(h, ptr, 0);
    }
    std::vector<double> buf(cells * 27);
    for (int s = 0; s < 4 && ok && st == EK_OK; ++s) {
        ok = fread(buf.data(), sizeof(double), cells * 27, f) == cells * 27;
        if (ok) st = ek_set_populations(h, s, buf.data());
    }
    fclose(f);
    if (st != EK_OK) return st;
    if (!ok) { ek_set_error(h, std::string(path) + ": truncated checkpoint"); return EK_ERR_INVALID; }
    st = ek_populations_restored(h);
    if (st != EK_OK) return st;
    h->steps = hd.steps;
    if (time) *time = hd.time;
    return EK_OK;
}

}  // extern "C"
