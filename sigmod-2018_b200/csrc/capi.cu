// capi.cu — Part 2 of include/b200_join.h: lifecycle, relation registration,
// and the kernel-level / fused entry points the parity tests and bench.py call
// through ctypes.  Plain pointers and sizes only.
#include "../../include/b200_join.h"
#include "engine.cuh"

#include <cmath>
#include <cstring>
#include <vector>

using namespace b200;

namespace {

// temporary device copy of a host uint64 array
DevBufPtr upload_u64(const uint64_t *host, uint64_t n) {
    Context  &c = ctx();
    DevBufPtr d = dev_alloc(n * sizeof(uint64_t));
    if (n) B200_CUDA(cudaMemcpyAsync(d->ptr, host, n * sizeof(uint64_t), cudaMemcpyHostToDevice, c.stream));
    return d;
}

// host uint64 ids -> device uint32 ids
DevBufPtr upload_ids(const uint64_t *host, uint64_t n) {
    DevBufPtr wide = upload_u64(host, n);
    DevBufPtr ids  = dev_alloc(n * sizeof(uint32_t));
    narrow_ids(wide->as<uint64_t>(), n, ids->as<uint32_t>());
    return ids;
}

void download_ids(const uint32_t *d, uint64_t n, uint64_t *out) {
    if (n == 0) return;
    Context  &c   = ctx();
    DevBufPtr tmp = dev_alloc(n * sizeof(uint64_t));
    widen_ids(d, n, tmp->as<uint64_t>());
    B200_CUDA(cudaMemcpyAsync(out, tmp->ptr, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c.stream));
    B200_CUDA(cudaStreamSynchronize(c.stream));
}

uint64_t host_max(const uint64_t *v, uint64_t n) {
    uint64_t m = 0;
    for (uint64_t i = 0; i < n; ++i) m = v[i] > m ? v[i] : m;
    return m;
}

int fail(const char *msg) {
    set_last_error(msg);
    return 1;
}

}  // namespace

extern "C" {

int b200_init(int device) {
    request_device(device);
    ensure_init();
    (void)ctx();
    return 0;
}

// the calling thread works on `device` from now on (its own context: stream, scratch); -1 = the process default
int b200_set_thread_device(int device) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device >= count) {
        cudaGetLastError();
        return fail("no such CUDA device");
    }
    ensure_init();
    set_thread_device(device);
    (void)ctx();
    return 0;
}

void b200_shutdown(void) {
    cudaDeviceSynchronize();
    unregister_all_columns();
}

const char *b200_last_error(void) { return last_error_string().c_str(); }

int b200_is_cuda(void) { return 1; }

int b200_register_relations(const relation_map *map, int count) {
    for (int r = 0; r < count; ++r)
        for (uint64_t j = 0; j < map[r].num_columns; ++j)
            register_host_column(map[r].columns[j], map[r].num_tuples, false);
    return 0;
}

// relation_map.c:53-83 on the GPU: fills map[r].col_stats[j] = {l, u, f, d} from the device copies (uploading the
// columns first when they are not registered yet)
int b200_compute_column_stats(relation_map *map, int count) {
    for (int r = 0; r < count; ++r) {
        if (!map[r].col_stats) continue;
        for (uint64_t j = 0; j < map[r].num_columns; ++j) {
            DevColumn     c  = lookup_column(map[r].columns[j], map[r].num_tuples);
            column_stats &st = map[r].col_stats[j];
            st.f             = (double)map[r].num_tuples;
            if (c.n && c.distinct == 0) {   // registered without statistics: compute them now
                uint64_t mx = 0;
                device_column_stats(c.d, c.n, &c.min_val, &mx, &c.distinct);
                c.max_val = mx;
            }
            st.l = c.min_val;
            st.u = c.n ? c.max_val : 0;
            st.d = (double)c.distinct;
        }
    }
    return 0;
}

int b200_register_device_column(const uint64_t *host_key, const uint64_t *device_ptr, uint64_t n,
                                uint64_t max_value) {
    register_device_column(host_key, device_ptr, n, max_value);
    return 0;
}

int b200_upload_column(const uint64_t *host_col, uint64_t n) {
    register_host_column(host_col, n, true);
    return 0;
}

void b200_unregister_all(void) { unregister_all_columns(); }

int b200_unregister_relations(const relation_map *map, int count) {
    for (int r = 0; r < count; ++r)
        for (uint64_t j = 0; j < map[r].num_columns; ++j) unregister_column(map[r].columns[j]);
    return 0;
}

// device memory for callers that keep relations resident in HBM (location = 1)
void *b200_device_malloc(uint64_t bytes) {
    ensure_init();
    (void)ctx();
    void *p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) {
        cudaGetLastError();
        set_last_error("cudaMalloc failed");
        return nullptr;
    }
    return p;
}

// Warm the stream-ordered pool: one allocation of `bytes` that goes straight back to the pool (which never returns
// memory to the driver, engine.cu configure_device_pool).  Physical allocations cost the driver 5-100 ms each on a
// cold process (b200 slow-log, profiles/r2_config5_x100_per_query.txt); a host does this once in its preparation
// phase so that no query pays for them.  Returns the bytes reserved (less when the device has less free memory).
uint64_t b200_reserve_device_memory(uint64_t bytes) {
    Context &c = ctx();
    size_t   free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    const uint64_t room = free_b > (4ull << 30) ? free_b - (4ull << 30) : 0;   // leave the driver some air
    if (bytes > room) bytes = room;
    if (bytes == 0) return 0;
    void *p = nullptr;
    if (cudaMallocAsync(&p, bytes, c.stream) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    B200_CUDA(cudaFreeAsync(p, c.stream));
    B200_CUDA(cudaStreamSynchronize(c.stream));
    return bytes;
}

void b200_device_free(void *device_ptr) {
    if (device_ptr) cudaFree(device_ptr);
}

int b200_copy_to_device(void *device_dst, const void *host_src, uint64_t bytes) {
    Context &c = ctx();
    B200_CUDA(cudaMemcpyAsync(device_dst, host_src, bytes, cudaMemcpyHostToDevice, c.stream));
    B200_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
}

int b200_copy_to_host(void *host_dst, const void *device_src, uint64_t bytes) {
    Context &c = ctx();
    B200_CUDA(cudaMemcpyAsync(host_dst, device_src, bytes, cudaMemcpyDeviceToHost, c.stream));
    B200_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
}

void *b200_get_stream(void) { return ctx().stream; }

int b200_set_stream(void *cuda_stream) {
    // the context's own stream stays alive (buffers allocated on it are freed in its order).  NULL is a stream
    // too: the legacy default stream (what torch.cuda.current_stream() is unless the caller changed it)
    Context &c = ctx();
    if (c.stream == c.own_stream) cudaStreamSynchronize(c.own_stream);   // adopted streams are the caller's to order
    c.stream = static_cast<cudaStream_t>(cuda_stream);
    return 0;
}

int b200_synchronize(void) {
    B200_CUDA(cudaStreamSynchronize(ctx().stream));
    return 0;
}

int b200_set_tuning(int radix_bits, int force_key64) {
    ensure_init();
    tuning().radix_bits  = radix_bits;
    tuning().force_key64 = force_key64;
    return 0;
}

int b200_set_profiling(int on) {
    set_profiling(on != 0);
    // (re-)enabling forgets the calling thread's timers: what b200_last_kernel_ms reports afterwards was recorded since
    if (on)
        for (auto &kv : ctx().timers) {
            kv.second.used       = false;
            kv.second.earlier_ms = 0;
            kv.second.scopes     = 0;
        }
    return 0;
}

// every scope of that name since profiling was (re-)enabled, not just the last one
double b200_sum_kernel_ms(const char *name, int *out_scopes) {
    Context &c  = ctx();
    auto     it = c.timers.find(name);
    if (out_scopes) *out_scopes = 0;
    if (it == c.timers.end() || !it->second.used) return -1.0;
    const double last = b200_last_kernel_ms(name);
    if (out_scopes) *out_scopes = it->second.scopes;
    return it->second.earlier_ms + (last > 0 ? last : 0.0);
}

double b200_last_kernel_ms(const char *name) {
    Context &c  = ctx();
    auto     it = c.timers.find(name);
    if (it == c.timers.end() || !it->second.used) return -1.0;
    float ms = 0.f;
    if (cudaEventSynchronize(it->second.stop) != cudaSuccess) return -1.0;
    if (cudaEventElapsedTime(&ms, it->second.start, it->second.stop) != cudaSuccess) return -1.0;
    return (double)ms;
}

uint64_t b200_kernel_launches(int reset) {
    uint64_t v = g_launches.load();
    if (reset) g_launches.store(0);
    return v;
}

// synthetic column straight into a DEVICE buffer (include/b200_synth.h)
int b200_synth_column(uint64_t *device_out, uint64_t first, uint64_t n, int kind, uint64_t k, uint64_t seed) {
    run_synth_column(device_out, first, n, kind, k, seed);
    B200_CUDA(cudaStreamSynchronize(ctx().stream));
    return 0;
}

// K1 (filter.c:115-170)
int b200_scan_filter(const uint64_t *col, uint64_t n, const uint64_t *ids, uint64_t n_ids, char cmp, int value,
                     uint64_t *out, uint64_t *out_n) {
    if (n > kMaxRows || n_ids > kMaxRows) return fail("more than 2^32-1 rows");
    if (cmp != '<' && cmp != '>' && cmp != '=') return fail("comparator must be <, > or =");
    DevBufPtr d_col = upload_u64(col, n);
    DevBufPtr d_ids = ids ? upload_ids(ids, n_ids) : nullptr;
    KeySrc    src{d_col->as<uint64_t>(), d_ids ? d_ids->as<uint32_t>() : nullptr,
               (uint32_t)(ids ? n_ids : n)};
    IdList l = run_filter(src, cmp, value);
    download_ids(l.ids->as<uint32_t>(), l.n, out);
    *out_n = l.n;
    return 0;
}

// K3-K5 (preprocess.c:13-178)
int b200_radix_partition(const uint64_t *keys, uint64_t n, int radix_bits, uint64_t *out_keys,
                         uint64_t *out_row_ids, uint64_t *out_hist, int64_t *out_psum) {
    if (n > kMaxRows) return fail("more than 2^32-1 rows");
    if (radix_bits < 0 || radix_bits > tuning().max_bits) return fail("radix_bits out of range");
    Context  &c = ctx();
    DevBufPtr d = upload_u64(keys, n);
    KeyVec    kv;
    kv.src     = KeySrc{d->as<uint64_t>(), nullptr, (uint32_t)n};
    kv.max_val = host_max(keys, n);
    PartitionOut   p      = run_partition(kv, radix_bits);
    const uint32_t nparts = 1u << radix_bits;
    std::vector<uint32_t> hist(nparts);
    B200_CUDA(cudaMemcpyAsync(hist.data(), p.hist->ptr, nparts * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                              c.stream));
    if (n) {
        DevBufPtr k64 = dev_alloc(n * sizeof(uint64_t)), r64 = dev_alloc(n * sizeof(uint64_t));
        unpack_partition(p, n, k64->as<uint64_t>(), r64->as<uint64_t>());
        B200_CUDA(cudaMemcpyAsync(out_keys, k64->ptr, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c.stream));
        B200_CUDA(cudaMemcpyAsync(out_row_ids, r64->ptr, n * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                                  c.stream));
    }
    B200_CUDA(cudaStreamSynchronize(c.stream));
    // preprocess.c:83-102: psum is the running start, -1 for an empty bucket
    int64_t run = 0;
    for (uint32_t b = 0; b < nparts; ++b) {
        out_hist[b] = hist[b];
        out_psum[b] = hist[b] ? run : -1;
        run += hist[b];
    }
    return 0;
}

// K6-K7 (rhjoin.c:13-111)
int b200_hash_join_pairs(const uint64_t *keys_r, uint64_t n_r, const uint64_t *keys_s, uint64_t n_s,
                         uint64_t *out_r, uint64_t *out_s, uint64_t cap, uint64_t *out_m) {
    if (n_r > kMaxRows || n_s > kMaxRows) return fail("more than 2^32-1 rows");
    *out_m = 0;
    if (n_r == 0 || n_s == 0) return 0;   // rhjoin.c:15-16
    DevBufPtr dr = upload_u64(keys_r, n_r), ds = upload_u64(keys_s, n_s);
    KeyVec    R, S;
    R.src     = KeySrc{dr->as<uint64_t>(), nullptr, (uint32_t)n_r};
    S.src     = KeySrc{ds->as<uint64_t>(), nullptr, (uint32_t)n_s};
    R.max_val = host_max(keys_r, n_r);
    S.max_val = host_max(keys_s, n_s);
    JoinResult j = run_join(R, S, JoinOut::Pairs, 0, nullptr);
    *out_m       = j.m;
    const uint64_t take = j.m < cap ? j.m : cap;
    download_ids(j.r_ids->as<uint32_t>(), take, out_r);
    download_ids(j.s_ids->as<uint32_t>(), take, out_s);
    return 0;
}

// K9 (inter_res.c:332-333)
int b200_gather_sum(const uint64_t *col, uint64_t n, const uint64_t *ids, uint64_t m, uint64_t *out_sum) {
    if (n > kMaxRows || m > kMaxRows) return fail("more than 2^32-1 rows");
    DevBufPtr       d_col = upload_u64(col, n);
    DevBufPtr       d_ids = upload_ids(ids, m);
    const uint64_t *cols[1] = {d_col->as<uint64_t>()};
    const uint32_t *idp[1]  = {d_ids->as<uint32_t>()};
    run_checksum(m, 1, cols, idp, out_sum);
    return 0;
}

// fused join -> SUM (rhjoin.c:13-111 folded into inter_res.c:320-339)
int b200_join_sum(const uint64_t *keys_r, uint64_t n_r, const uint64_t *keys_s, uint64_t n_s, uint64_t max_key,
                  int n_proj, const uint64_t *const *proj, const int *proj_side, int location,
                  uint64_t *out_sums, uint64_t *out_matches) {
    if (n_r > kMaxRows || n_s > kMaxRows) return fail("more than 2^32-1 rows");
    if (n_proj < 0 || n_proj > kMaxProj) return fail("at most 8 fused projections");
    for (int k = 0; k < n_proj; ++k) out_sums[k] = 0;
    *out_matches = 0;
    if (n_r == 0 || n_s == 0) return 0;
    std::vector<DevBufPtr> keep;
    KeyVec                 R, S;
    ProjDesc               pd[kMaxProj];
    if (location == 0) {
        // end-to-end arm: every input crosses PCIe inside the call
        DevBufPtr dr = upload_u64(keys_r, n_r), ds = upload_u64(keys_s, n_s);
        keep.push_back(dr);
        keep.push_back(ds);
        R.src = KeySrc{dr->as<uint64_t>(), nullptr, (uint32_t)n_r};
        S.src = KeySrc{ds->as<uint64_t>(), nullptr, (uint32_t)n_s};
        for (int k = 0; k < n_proj; ++k) {
            DevBufPtr dp = upload_u64(proj[k], proj_side[k] == 0 ? n_r : n_s);
            keep.push_back(dp);
            pd[k] = ProjDesc{dp->as<uint64_t>(), nullptr, proj_side[k], nullptr};
        }
    } else {
        R.src = KeySrc{keys_r, nullptr, (uint32_t)n_r};
        S.src = KeySrc{keys_s, nullptr, (uint32_t)n_s};
        for (int k = 0; k < n_proj; ++k) pd[k] = ProjDesc{proj[k], nullptr, proj_side[k], nullptr};
    }
    R.max_val = S.max_val = max_key;
    JoinResult j = run_join(R, S, JoinOut::Sum, n_proj, pd);
    for (int k = 0; k < n_proj; ++k) out_sums[k] = j.sums[k];
    *out_matches = j.m;
    return 0;
}

// ---- staged join + CUDA IPC (multi-GPU, one process per GPU) ---------------
int b200_ipc_export(const void *device_ptr, unsigned char *out_handle64) {
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, const_cast<void *>(device_ptr)) != cudaSuccess) {
        cudaGetLastError();
        return fail("cudaIpcGetMemHandle failed (the buffer must come from b200_device_malloc)");
    }
    static_assert(sizeof(h) == 64, "IPC handle size");
    memcpy(out_handle64, &h, 64);
    return 0;
}

void *b200_ipc_import(const unsigned char *handle64) {
    ensure_init();
    (void)ctx();
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void *p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        set_last_error("cudaIpcOpenMemHandle failed (no peer access between the two GPUs?)");
        return nullptr;
    }
    return p;
}

int b200_ipc_close(void *imported_ptr) {
    return cudaIpcCloseMemHandle(imported_ptr) == cudaSuccess ? 0 : fail("cudaIpcCloseMemHandle failed");
}

int b200_stage_hist(const uint64_t *d_keys, uint64_t n, int radix_bits, uint32_t *d_hist) {
    if (n > kMaxRows) return fail("more than 2^32-1 rows");
    stage_hist(d_keys, n, radix_bits, d_hist);
    return 0;
}

int b200_stage_scatter_build(const uint64_t *d_keys, uint64_t n, uint32_t rid_base, int radix_bits,
                             const uint32_t *d_hist_local, const uint32_t *d_dst_start, int ndst,
                             void *const *tup_dst, int npay, const uint64_t *const *pay_cols,
                             uint64_t *const *pay_dst, int phase) {
    if (n > kMaxRows) return fail("more than 2^32-1 rows");
    if (ndst < 1 || ndst > 8 || npay < 0 || npay > 2) return fail("ndst must be 1..8 and npay 0..2");
    if (phase < 0 || phase > 2) return fail("phase must be 0, 1 or 2");
    stage_scatter_build(d_keys, n, rid_base, radix_bits, d_hist_local, d_dst_start, ndst, tup_dst, npay, pay_cols,
                        pay_dst, phase);
    return 0;
}

int b200_stage_scatter_probe(const uint64_t *d_keys, uint64_t n, int radix_bits, uint32_t *d_cursor, void *d_tup_out) {
    if (n > kMaxRows) return fail("more than 2^32-1 rows");
    stage_scatter_probe(d_keys, n, radix_bits, d_cursor, d_tup_out);
    return 0;
}

uint32_t b200_opt_region_cap(uint64_t n_probe, int radix_bits) { return opt_region_cap(n_probe, radix_bits); }

int b200_stage_build_cursors(const uint32_t *d_hist_all, int world, int rank, int radix_bits, uint32_t *d_total,
                             uint32_t *d_my_start) {
    if (world < 1 || world > 8 || rank < 0 || rank >= world) return fail("bad world / rank");
    stage_build_cursors(d_hist_all, world, rank, radix_bits, d_total, d_my_start);
    return 0;
}

int b200_stage_join_sum_async(const void *d_tup_b, const uint32_t *d_hist_b, const void *d_tup_p,
                              const uint32_t *d_hist_p, int radix_bits, int n_proj, const uint64_t *const *proj_cols,
                              const int *proj_side, const uint64_t *const *proj_part_vals, uint32_t opt_cap,
                              const void *d_ov, const uint32_t *d_ovcnt, uint64_t *d_result) {
    if (n_proj < 0 || n_proj > kMaxProj) return fail("at most 8 fused projections");
    ProjDesc pd[kMaxProj];
    for (int k = 0; k < n_proj; ++k)
        pd[k] = ProjDesc{proj_cols[k], nullptr, proj_side[k], proj_part_vals ? proj_part_vals[k] : nullptr};
    stage_join_sum(d_tup_b, d_hist_b, d_tup_p, d_hist_p, radix_bits, n_proj, pd, opt_cap, d_ov, d_ovcnt,
                   reinterpret_cast<unsigned long long *>(d_result));
    return 0;
}

int b200_stage_scatter_probe_opt(const uint64_t *d_keys, uint64_t n, int radix_bits, uint32_t opt_cap,
                                 uint32_t *d_cursor, void *d_tup_out, void *d_ov, uint32_t *d_ovcnt) {
    if (n > (1u << 30)) return fail("histogram-free scatter is limited to 2^30 probe rows");
    stage_scatter_probe_opt(d_keys, n, radix_bits, opt_cap, d_cursor, d_tup_out, d_ov, d_ovcnt);
    return 0;
}

int b200_stage_scatter_probe_opt_carry(const uint64_t *d_keys, uint64_t n, int radix_bits, uint32_t opt_cap,
                                       uint32_t *d_cursor, void *d_tup_out, void *d_ov, uint32_t *d_ovcnt,
                                       const uint64_t *d_carry_col) {
    if (n > (1u << 30)) return fail("histogram-free scatter is limited to 2^30 probe rows");
    if (!d_carry_col || (reinterpret_cast<uintptr_t>(d_carry_col) & 7)) return fail("bad carried column");
    stage_scatter_probe_opt(d_keys, n, radix_bits, opt_cap, d_cursor, d_tup_out, d_ov, d_ovcnt, d_carry_col);
    return 0;
}

int b200_stage_join_sum(const void *d_tup_b, const uint32_t *d_hist_b, const void *d_tup_p, const uint32_t *d_hist_p,
                        int radix_bits, int n_proj, const uint64_t *const *proj_cols, const int *proj_side,
                        const uint64_t *const *proj_part_vals, uint32_t opt_cap, const void *d_ov,
                        const uint32_t *d_ovcnt, uint64_t *out_sums, uint64_t *out_matches) {
    if (n_proj < 0 || n_proj > kMaxProj) return fail("at most 8 fused projections");
    ProjDesc pd[kMaxProj];
    for (int k = 0; k < n_proj; ++k)
        pd[k] = ProjDesc{proj_cols[k], nullptr, proj_side[k], proj_part_vals ? proj_part_vals[k] : nullptr};
    JoinResult j = stage_join_sum(d_tup_b, d_hist_b, d_tup_p, d_hist_p, radix_bits, n_proj, pd, opt_cap, d_ov, d_ovcnt);
    for (int k = 0; k < n_proj; ++k) out_sums[k] = j.sums[k];
    *out_matches = j.m;
    return 0;
}

int b200_stage_scatter_build_local(const uint64_t *d_keys, uint64_t n, uint32_t rid_base, int radix_bits,
                                   const uint32_t *d_hist_local, void *d_tup_out, int npay,
                                   const uint64_t *const *pay_cols, uint64_t *const *pay_out) {
    if (n > kMaxRows) return fail("more than 2^32-1 rows");
    if (npay < 0 || npay > 2) return fail("npay must be 0..2");
    stage_scatter_build_local(d_keys, n, rid_base, radix_bits, d_hist_local, d_tup_out, npay, pay_cols, pay_out);
    return 0;
}

int b200_copy_device_async(void *dst, const void *src, uint64_t bytes) {
    B200_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx().stream));
    return 0;
}

int b200_stage_join_sum_seg(const void *d_tup_b, const uint32_t *d_hist_all, int nseg, uint32_t seg_rows,
                            const void *d_tup_p, const uint32_t *d_hist_p, int radix_bits, int n_proj,
                            const uint64_t *const *proj_cols, const int *proj_side,
                            const uint64_t *const *proj_part_vals, uint32_t opt_cap, const void *d_ov,
                            const uint32_t *d_ovcnt, uint64_t *d_result, uint64_t *out_sums, uint64_t *out_matches) {
    if (n_proj < 0 || n_proj > kMaxProj) return fail("at most 8 fused projections");
    if (nseg < 1 || nseg > 8) return fail("nseg must be 1..8");
    ProjDesc pd[kMaxProj];
    for (int k = 0; k < n_proj; ++k)
        pd[k] = ProjDesc{proj_cols[k], nullptr, proj_side[k], proj_part_vals ? proj_part_vals[k] : nullptr};
    JoinResult j = stage_join_sum(d_tup_b, d_hist_all, d_tup_p, d_hist_p, radix_bits, n_proj, pd, opt_cap, d_ov, d_ovcnt,
                                  reinterpret_cast<unsigned long long *>(d_result), nseg, seg_rows);
    if (!d_result) {
        for (int k = 0; k < n_proj; ++k) out_sums[k] = j.sums[k];
        *out_matches = j.m;
    }
    return 0;
}

int b200_stage_exchange_cursors(const uint32_t *d_hist_all, int world, int rank, int radix_bits, uint32_t cap,
                                uint32_t *d_src_off, uint32_t *d_dst_start, uint32_t *d_own_total, uint32_t *d_need) {
    if (world < 1 || world > 8 || rank < 0 || rank >= world) return fail("bad world / rank");
    if (radix_bits < 2 || (1u << radix_bits) < (uint32_t)world) return fail("fewer partitions than ranks");
    stage_exchange_cursors(d_hist_all, world, rank, radix_bits, cap, d_src_off, d_dst_start, d_own_total, d_need);
    return 0;
}

int b200_stage_exchange_segments(const void *d_src_tup, int npay, const uint64_t *const *src_pay, uint64_t n,
                                 int radix_bits, int world, const uint32_t *d_src_off, const uint32_t *d_dst_start,
                                 uint32_t cap, int rewrite_rid, void *const *tup_dst, uint64_t *const *pay_dst) {
    if (world < 1 || world > 8) return fail("bad world");
    if (npay < 0 || npay > 2) return fail("npay must be 0..2");
    if (n > kMaxRows) return fail("more than 2^32-1 rows");
    stage_exchange_segments(d_src_tup, npay, src_pay, n, radix_bits, world, d_src_off, d_dst_start, cap, rewrite_rid,
                            tup_dst, pay_dst);
    return 0;
}

int b200_radix_bits_for(uint64_t n_build) {
    // the library's automatic choice for a 32-bit-key build side of n_build rows
    return auto_radix_bits(n_build, false);
}

}  // extern "C"
