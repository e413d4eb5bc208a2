// multi.inl — one host thread, several GPUs, behind the C ABI (vpmb200_multi_*; include/vpmb200.h).  Included at the end of
// engine.cu: it drives ordinary per-device engines through their internal entry points.
//
// The reference's host (Julia) is ONE process making blocking calls on one ParticleField
// (/root/reference/src/FLOWUnsteady_simulation.jl:339-447), so the multi-GPU field must be reachable from one thread through
// the same upload / uj / nextstep / add / remove / download calls.  A vpmb200_multi owns one engine per device; particles are
// sharded, the GLOBAL particle order the host sees (the order of its particle matrix, on which
// vpm.remove_particle's swap-with-last semantics are defined) is a host-side map gid <-> (shard, local slot).
//   * direct U/J (+E_str): every device packs its source tiles, the tile sets are exchanged device-to-device with
//     cudaMemcpyPeerAsync (NVLink / NVSwitch; no host staging, no NCCL), each device's pair kernel runs on its own tiles first
//     and then on each peer's set; cross-device ordering by CUDA events, all work is enqueued asynchronously so the devices
//     run concurrently;
//   * per-particle stages (SFS coefficient, update, relaxation) are shard-local;
//   * add_particle appends to the least-loaded shard, remove_particle / remove_where reproduce the reference's resulting
//     GLOBAL order through the map, vpmb200_multi_rebalance moves the tail of the fullest shard to the emptiest one.
//   * UJ_fmm: the local-essential-tree phases of fmm_let.cuh (vpmb200_let_*) with the exchanges done by peer copies and small
//     host-side reductions — the sequence flowunsteady_b200/dist.py runs across processes with NCCL, here under one thread
//     (phases separated by stream synchronisation; no far-field reuse between DynamicSFS's two filter evaluations yet).

struct vpmb200_multi {
    int G = 0;
    std::vector<vpmb200_engine*> eng;
    std::vector<int> dev;
    int64_t maxp = 0, np = 0, cap_local = 0;
    std::vector<std::vector<int64_t>> gid;   // gid[k][l]: global index of shard k's local slot l
    std::vector<int32_t> owner;              // by global index
    std::vector<int64_t> local;
    std::vector<double*> tiles_own;          // per device: its packed source tiles
    std::vector<double*> tiles_peer;         // per device: room for every peer's tile set (slot_tiles each)
    int64_t slot_tiles = 0;
    std::vector<cudaEvent_t> ev_pack, ev_done;
    bool have_done = false;
    std::vector<double> stage;               // host staging for non-contiguous transfers
    // UJ_fmm with a local essential tree under one host thread: per-device exchange buffers (grow-only)
    struct LetBuf { void* p = nullptr; size_t bytes = 0; };
    std::vector<LetBuf> lb_send, lb_recv, lb_cells, lb_M, lb_rec, lb_out, lb_res;
    bool let_work_ready = false;
    // static-particle fast path on the sharded field (vpmb200_multi_set_statics): the step's statics are ordinary appended particles (they sit at the END
    // of the global order, so dropping them is a pop) that nextstep removes again; valid while nt == static_gen
    int64_t nstatic = 0, static_gen = -1;
    vpmb200_schemes sch;
    double t = 0.0;
    int64_t nt = 0;
    std::string err;
};

namespace {

thread_local std::string g_multi_create_error;

int32_t mfail(vpmb200_multi* m, int32_t code, const std::string& msg) {
    if (m) m->err = msg; else g_multi_create_error = msg;
    return code;
}

// error of a per-device engine call -> the multi handle's message
#define M_ENG(m, k, call)                                                                                   \
    do {                                                                                                    \
        int32_t _rc = (call);                                                                               \
        if (_rc != VPMB200_OK)                                                                              \
            return mfail((m), _rc, "GPU " + std::to_string((m)->dev[k]) + ": " + vpmb200_last_error((m)->eng[k])); \
    } while (0)
#define M_CU(m, call)                                                                                       \
    do {                                                                                                    \
        cudaError_t _st = (call);                                                                           \
        if (_st != cudaSuccess) return mfail((m), VPMB200_ECUDA, std::string(#call) + ": " + cudaGetErrorString(_st)); \
    } while (0)

bool multi_is_blocked(const vpmb200_multi* m) {   // shard k holds the consecutive global indices [off_k, off_k + n_k) in order
    int64_t off = 0;
    for (int k = 0; k < m->G; ++k) {
        const auto& g = m->gid[k];
        for (size_t l = 0; l < g.size(); ++l)
            if (g[l] != off + (int64_t)l) return false;
        off += (int64_t)g.size();
    }
    return true;
}

void multi_rebuild_inverse(vpmb200_multi* m) {
    m->owner.assign((size_t)m->np, 0);
    m->local.assign((size_t)m->np, 0);
    for (int k = 0; k < m->G; ++k)
        for (size_t l = 0; l < m->gid[k].size(); ++l) {
            m->owner[(size_t)m->gid[k][l]] = k;
            m->local[(size_t)m->gid[k][l]] = (int64_t)l;
        }
}

// The arrangement the reference's removal loop leaves behind (i = n..1: if particle i goes, the current last particle moves
// into slot i; src/FLOWUnsteady_processing.jl:50-187): returns a[pos] = original index of the particle that ends at pos.
std::vector<int64_t> reference_loop_order(const std::vector<char>& keep) {
    const int64_t n = (int64_t)keep.size();
    std::vector<int64_t> a((size_t)n);
    for (int64_t i = 0; i < n; ++i) a[(size_t)i] = i;
    int64_t last = n - 1;
    for (int64_t i = n - 1; i >= 0; --i)
        if (!keep[(size_t)i]) {          // position i still holds original particle i when the cursor reaches it
            a[(size_t)i] = a[(size_t)last];
            --last;
        }
    a.resize((size_t)(last + 1));
    return a;
}

int32_t multi_wait_peers_done(vpmb200_multi* m) {
    // nobody may overwrite its tile buffer while a peer is still copying the previous set out of it
    if (!m->have_done) return VPMB200_OK;
    for (int k = 0; k < m->G; ++k) {
        M_CU(m, cudaSetDevice(m->dev[k]));
        for (int q = 0; q < m->G; ++q)
            if (q != k) M_CU(m, cudaStreamWaitEvent(m->eng[k]->stream, m->ev_done[q], 0));
    }
    return VPMB200_OK;
}

// pack on every device -> peer copies -> pair kernel on own tiles, then on every peer's set
template <typename Pack, typename Apply>
int32_t multi_pairwise(vpmb200_multi* m, Pack pack, Apply apply) {
    int32_t rc = multi_wait_peers_done(m);
    if (rc) return rc;
    std::vector<int64_t> ntiles(m->G);
    for (int k = 0; k < m->G; ++k) {
        ntiles[k] = vpmb200_tiles_for(m->eng[k]->np);
        M_ENG(m, k, pack(m->eng[k], m->tiles_own[k]));
        M_CU(m, cudaSetDevice(m->dev[k]));
        M_CU(m, cudaEventRecord(m->ev_pack[k], m->eng[k]->stream));
    }
    for (int k = 0; k < m->G; ++k) {
        if (m->eng[k]->np > 0) M_ENG(m, k, apply(m->eng[k], m->tiles_own[k], ntiles[k]));
        M_CU(m, cudaSetDevice(m->dev[k]));
        for (int q = 0; q < m->G; ++q) {
            if (q == k || ntiles[q] == 0) continue;
            double* dst = m->tiles_peer[k] + (size_t)q * m->slot_tiles * TILE_DOUBLES;
            M_CU(m, cudaStreamWaitEvent(m->eng[k]->stream, m->ev_pack[q], 0));
            M_CU(m, cudaMemcpyPeerAsync(dst, m->dev[k], m->tiles_own[q], m->dev[q], sizeof(double) * (size_t)ntiles[q] * TILE_DOUBLES,
                                        m->eng[k]->stream));
        }
        M_CU(m, cudaEventRecord(m->ev_done[k], m->eng[k]->stream));
        if (m->eng[k]->np > 0)
            for (int q = 0; q < m->G; ++q) {
                if (q == k || ntiles[q] == 0) continue;
                M_ENG(m, k, apply(m->eng[k], m->tiles_peer[k] + (size_t)q * m->slot_tiles * TILE_DOUBLES, ntiles[q]));
            }
    }
    m->have_done = true;
    return VPMB200_OK;
}

// ---- UJ_fmm with a local essential tree, one host thread: the phases of fmm_let.cuh with the exchanges done by peer copies.
//      Same sequence as flowunsteady_b200/dist.py (_uj_fmm_let); phases are separated by stream synchronisation.
int32_t let_grow(vpmb200_multi* m, std::vector<vpmb200_multi::LetBuf>& v, int k, size_t bytes) {
    if ((int)v.size() < m->G) v.resize(m->G);
    if (bytes <= v[k].bytes && v[k].p) return VPMB200_OK;
    M_CU(m, cudaSetDevice(m->dev[k]));
    if (v[k].p) cudaFree(v[k].p);
    v[k].p = nullptr;
    v[k].bytes = 0;
    const size_t cap = bytes + bytes / 8 + 4096;
    M_CU(m, cudaMalloc(&v[k].p, cap));
    v[k].bytes = cap;
    return VPMB200_OK;
}

int32_t multi_sync_all(vpmb200_multi* m) {
    for (int k = 0; k < m->G; ++k) M_ENG(m, k, vpmb200_synchronize(m->eng[k]));
    return VPMB200_OK;
}

// every device's `bins` values of type T combined on the host (sum or max) and written back to every device
template <typename T, typename Op>
int32_t multi_allreduce(vpmb200_multi* m, const std::vector<T*>& dev_ptrs, int count, Op op) {
    std::vector<T> acc(count), tmp(count);
    for (int k = 0; k < m->G; ++k) {
        M_CU(m, cudaSetDevice(m->dev[k]));
        M_CU(m, cudaMemcpy(k == 0 ? acc.data() : tmp.data(), dev_ptrs[k], sizeof(T) * count, cudaMemcpyDeviceToHost));
        if (k > 0)
            for (int i = 0; i < count; ++i) acc[i] = op(acc[i], tmp[i]);
    }
    for (int k = 0; k < m->G; ++k) {
        M_CU(m, cudaSetDevice(m->dev[k]));
        M_CU(m, cudaMemcpy(dev_ptrs[k], acc.data(), sizeof(T) * count, cudaMemcpyHostToDevice));
    }
    return VPMB200_OK;
}

// rows of `ncol` doubles: block (src k -> dst q) of src's buffer goes behind the blocks of the lower source ranks in dst's buffer
int32_t multi_alltoall_rows(vpmb200_multi* m, std::vector<vpmb200_multi::LetBuf>& src, std::vector<vpmb200_multi::LetBuf>& dst,
                            const std::vector<std::vector<int64_t>>& counts /* [src][dst] */, int ncol, bool transpose) {
    // transpose = false: src k sends counts[k][q] rows to q;  true (the way back): src q sends counts[k][q] rows to k
    const int G = m->G;
    for (int to = 0; to < G; ++to) {
        int64_t need = 0;
        for (int from = 0; from < G; ++from) need += transpose ? counts[to][from] : counts[from][to];
        int32_t rc = let_grow(m, dst, to, sizeof(double) * (size_t)std::max<int64_t>(need, 1) * ncol);
        if (rc) return rc;
    }
    for (int to = 0; to < G; ++to) {
        M_CU(m, cudaSetDevice(m->dev[to]));
        int64_t doff = 0;
        for (int from = 0; from < G; ++from) {
            const int64_t n = transpose ? counts[to][from] : counts[from][to];
            int64_t soff = 0;
            for (int q = 0; q < to; ++q) soff += transpose ? counts[q][from] : counts[from][q];
            if (n > 0)
                M_CU(m, cudaMemcpyPeerAsync(static_cast<double*>(dst[to].p) + doff * ncol, m->dev[to],
                                            static_cast<const double*>(src[from].p) + soff * ncol, m->dev[from],
                                            sizeof(double) * (size_t)n * ncol, m->eng[to]->stream));
            doff += n;
        }
    }
    return multi_sync_all(m);
}

int32_t multi_uj_fmm(vpmb200_multi* m, int reset, int reset_sfs, int sfs) {
    if (sfs && !reset) return mfail(m, VPMB200_ENOTSUP, "sharded UJ_fmm with sfs needs reset (what every SFS scheme calls)");
    const int G = m->G, Lc = 5, bins = 1 << (3 * Lc);
    // 1 bounds
    double lohi[6] = {1e300, 1e300, 1e300, -1e300, -1e300, -1e300};
    for (int k = 0; k < G; ++k) {
        double b[6];
        M_ENG(m, k, vpmb200_let_bounds(m->eng[k], b));
        for (int c = 0; c < 3; ++c) { lohi[c] = std::min(lohi[c], b[c]); lohi[3 + c] = std::max(lohi[3 + c], b[3 + c]); }
    }
    if (!(lohi[0] <= lohi[3])) return VPMB200_OK;   // no particle anywhere
    // 2 keys + histogram, summed over the devices
    std::vector<int*> hist(G);
    std::vector<double*> bmax(G);
    std::vector<long long*> work(G);
    bool nzs = false;
    for (int k = 0; k < G; ++k) {
        void *h = nullptr, *b = nullptr, *w = nullptr;
        M_ENG(m, k, vpmb200_let_keys(m->eng[k], lohi, Lc, &h, &b));
        M_ENG(m, k, vpmb200_let_work(m->eng[k], &w));
        hist[k] = static_cast<int*>(h);
        bmax[k] = static_cast<double*>(b);
        work[k] = static_cast<long long*>(w);
        nzs = b != nullptr;
    }
    int32_t rc = multi_sync_all(m);
    if (rc) return rc;
    if ((rc = multi_allreduce<int>(m, hist, bins, [](int a, int b) { return a + b; }))) return rc;
    if (nzs && (rc = multi_allreduce<double>(m, bmax, bins, [](double a, double b) { return std::max(a, b); }))) return rc;
    // 3 partition (the work counts of the previous evaluation were summed right after it)
    std::vector<std::vector<int64_t>> counts(G, std::vector<int64_t>(G, 0));   // [home k][owner q]
    for (int k = 0; k < G; ++k) M_ENG(m, k, vpmb200_let_partition(m->eng[k], G, k, m->let_work_ready ? 1 : 0, counts[k].data()));
    std::vector<int64_t> n_own(G, 0);
    int64_t n_all = 0;
    for (int k = 0; k < G; ++k)
        for (int q = 0; q < G; ++q) { n_own[q] += counts[k][q]; n_all += counts[k][q]; }
    // 4 particle rows to their owners
    for (int k = 0; k < G; ++k) {
        if ((rc = let_grow(m, m->lb_send, k, sizeof(double) * (size_t)std::max<int64_t>(m->eng[k]->np, 1) * LET_ROW))) return rc;
        M_ENG(m, k, vpmb200_let_pack(m->eng[k], static_cast<double*>(m->lb_send[k].p)));
    }
    if ((rc = multi_sync_all(m))) return rc;
    if ((rc = multi_alltoall_rows(m, m->lb_send, m->lb_recv, counts, LET_ROW, false))) return rc;
    // 5 owner trees + upward pass
    std::vector<int64_t> ncells(G, 0);
    int64_t nm3 = 0;
    for (int q = 0; q < G; ++q) {
        int64_t info[4];
        M_ENG(m, q, vpmb200_let_build(m->eng[q], static_cast<const double*>(m->lb_recv[q].p), n_own[q], n_all, 0, info));
        ncells[q] = info[0];
        nm3 = info[2];
    }
    if ((rc = multi_sync_all(m))) return rc;
    // 6 skeletons, multipoles, records of every rank gathered on every device (peer copies), appended behind the own ones
    const int64_t slot_c = std::max<int64_t>(*std::max_element(ncells.begin(), ncells.end()), 1);
    const int64_t slot_n = std::max<int64_t>(*std::max_element(n_own.begin(), n_own.end()), 1);
    const size_t cb = (size_t)vpmb200_let_cell_bytes();
    auto gather_records = [&]() -> int32_t {
        for (int q = 0; q < G; ++q) {
            int32_t r2 = let_grow(m, m->lb_rec, q, sizeof(double) * (size_t)G * slot_n * REC_REALS);
            if (r2) return r2;
            M_CU(m, cudaSetDevice(m->dev[q]));
            for (int k = 0; k < G; ++k) {
                if (k == q || n_own[k] == 0) continue;
                void* p3[3];
                vpmb200_let_ptrs(m->eng[k], p3);
                M_CU(m, cudaMemcpyPeerAsync(static_cast<double*>(m->lb_rec[q].p) + (size_t)k * slot_n * REC_REALS, m->dev[q], p3[2], m->dev[k],
                                            sizeof(double) * (size_t)n_own[k] * REC_REALS, m->eng[q]->stream));
            }
            M_ENG(m, q, vpmb200_let_attach_records(m->eng[q], static_cast<const double*>(m->lb_rec[q].p), slot_n, n_own.data()));
        }
        return multi_sync_all(m);
    };
    for (int q = 0; q < G; ++q) {
        if ((rc = let_grow(m, m->lb_cells, q, cb * (size_t)G * slot_c))) return rc;
        if ((rc = let_grow(m, m->lb_M, q, sizeof(double) * (size_t)G * slot_c * nm3))) return rc;
        M_CU(m, cudaSetDevice(m->dev[q]));
        for (int k = 0; k < G; ++k) {
            if (k == q || ncells[k] == 0) continue;
            void* p3[3];
            vpmb200_let_ptrs(m->eng[k], p3);
            M_CU(m, cudaMemcpyPeerAsync(static_cast<char*>(m->lb_cells[q].p) + (size_t)k * slot_c * cb, m->dev[q], p3[0], m->dev[k],
                                        cb * (size_t)ncells[k], m->eng[q]->stream));
            M_CU(m, cudaMemcpyPeerAsync(static_cast<double*>(m->lb_M[q].p) + (size_t)k * slot_c * nm3, m->dev[q], p3[1], m->dev[k],
                                        sizeof(double) * (size_t)ncells[k] * nm3, m->eng[q]->stream));
        }
        M_ENG(m, q, vpmb200_let_attach_tree(m->eng[q], m->lb_cells[q].p, static_cast<const double*>(m->lb_M[q].p), slot_c, ncells.data(),
                                            n_own.data()));
    }
    if ((rc = multi_sync_all(m))) return rc;
    if ((rc = gather_records())) return rc;
    // 7 evaluate, 8 results home
    for (int q = 0; q < G; ++q) {
        if ((rc = let_grow(m, m->lb_out, q, sizeof(double) * (size_t)std::max<int64_t>(n_own[q], 1) * 12))) return rc;
        M_ENG(m, q, vpmb200_let_evaluate(m->eng[q], static_cast<double*>(m->lb_out[q].p), 0, 0));
    }
    if ((rc = multi_sync_all(m))) return rc;
    {   // the interaction work counted per Morton bin, summed: the next cut equalises it
        if ((rc = multi_allreduce<long long>(m, work, bins, [](long long a, long long b) { return a + b; }))) return rc;
        m->let_work_ready = true;
    }
    if ((rc = multi_alltoall_rows(m, m->lb_out, m->lb_res, counts, 12, true))) return rc;
    for (int k = 0; k < G; ++k) {
        M_ENG(m, k, vpmb200_let_finish(m->eng[k], static_cast<const double*>(m->lb_res[k].p), 0, reset));
        if (reset_sfs) M_ENG(m, k, vpmb200_reset_particles_sfs(m->eng[k]));
    }
    if (sfs) {
        for (int q = 0; q < G; ++q) M_ENG(m, q, vpmb200_let_estr_records(m->eng[q]));
        if ((rc = multi_sync_all(m))) return rc;
        if ((rc = gather_records())) return rc;
        for (int q = 0; q < G; ++q) M_ENG(m, q, vpmb200_let_estr_evaluate(m->eng[q], static_cast<double*>(m->lb_out[q].p)));
        if ((rc = multi_sync_all(m))) return rc;
        if ((rc = multi_alltoall_rows(m, m->lb_out, m->lb_res, counts, 3, true))) return rc;
        for (int k = 0; k < G; ++k) M_ENG(m, k, vpmb200_let_finish(m->eng[k], static_cast<const double*>(m->lb_res[k].p), 1, 0));
    }
    return multi_sync_all(m);
}

int32_t multi_uj(vpmb200_multi* m, int reset, int reset_sfs, int sfs) {
    if (m->sch.uj == VPMB200_UJ_FMM) return multi_uj_fmm(m, reset, reset_sfs, sfs);
    for (int k = 0; k < m->G; ++k) {
        if (reset) M_ENG(m, k, vpmb200_reset_particles(m->eng[k]));
        if (reset_sfs) M_ENG(m, k, vpmb200_reset_particles_sfs(m->eng[k]));
    }
    int32_t rc = multi_pairwise(m, [](vpmb200_engine* e, double* d) { return vpmb200_pack_uj_records(e, d); },
                                [](vpmb200_engine* e, const double* t, int64_t n) { return vpmb200_uj_from_records(e, t, n, 1); });
    if (rc) return rc;
    if (sfs)
        rc = multi_pairwise(m, [](vpmb200_engine* e, double* d) { return vpmb200_pack_estr_records(e, d); },
                            [](vpmb200_engine* e, const double* t, int64_t n) { return vpmb200_estr_from_records(e, t, n); });
    return rc;
}

int32_t multi_stage(vpmb200_multi* m, int stage, double a = 0, double b = 0, double dt = 0, const double* Uinf = nullptr) {
    for (int k = 0; k < m->G; ++k) M_ENG(m, k, vpmb200_stage(m->eng[k], stage, a, b, dt, Uinf));
    return VPMB200_OK;
}

// pfield.SFS(pfield; a, b) over the shards (the sequence of do_sfs)
int32_t multi_sfs(vpmb200_multi* m, double a, double b) {
    (void)b;
    const bool first = (a == 1.0 || a == 0.0);
    int32_t rc;
    switch (m->sch.sfs) {
    case VPMB200_SFS_NONE:
        return multi_uj(m, 1, 0, 0);
    case VPMB200_SFS_CONSTANT:
        if ((rc = multi_uj(m, 1, 1, 1))) return rc;
        if (first) {
            if ((rc = multi_stage(m, VPMB200_STAGE_CONSTANT_COEFF))) return rc;
            if ((rc = multi_stage(m, VPMB200_STAGE_CLIP_CONTROL))) return rc;
        }
        return VPMB200_OK;
    default:
        if (!first) return multi_uj(m, 1, 1, 1);
        if ((rc = multi_stage(m, VPMB200_STAGE_SCALE_SIGMA_TEST))) return rc;
        if ((rc = multi_uj(m, 1, 1, 1))) return rc;
        if ((rc = multi_stage(m, VPMB200_STAGE_STORE_TEST))) return rc;
        if ((rc = multi_stage(m, VPMB200_STAGE_SCALE_SIGMA_DOMAIN))) return rc;
        if ((rc = multi_uj(m, 1, 1, 1))) return rc;
        if ((rc = multi_stage(m, VPMB200_STAGE_DYNAMIC_COEFF))) return rc;
        return multi_stage(m, VPMB200_STAGE_CLIP_CONTROL);
    }
}

// host rows of global indices gid[k][*] <-> shard k, through the staging buffer when they are not one contiguous run
int32_t multi_transfer(vpmb200_multi* m, double* particles, int64_t ld, uint32_t mask, bool up) {
    const bool blocked = multi_is_blocked(m);
    int64_t off = 0;
    for (int k = 0; k < m->G; ++k) {
        const int64_t n = (int64_t)m->gid[k].size();
        if (up) {
            if (blocked) {
                M_ENG(m, k, vpmb200_upload(m->eng[k], n ? particles + off * ld : nullptr, ld, n, mask));
            } else {
                m->stage.resize((size_t)std::max<int64_t>(n, 1) * NFIELDS);
                for (int64_t l = 0; l < n; ++l)
                    std::memcpy(&m->stage[(size_t)l * NFIELDS], particles + m->gid[k][(size_t)l] * ld, sizeof(double) * NFIELDS);
                M_ENG(m, k, vpmb200_upload(m->eng[k], n ? m->stage.data() : nullptr, NFIELDS, n, mask));
            }
        } else if (n > 0) {
            if (blocked) {
                M_ENG(m, k, vpmb200_download(m->eng[k], particles + off * ld, ld, n, mask));
            } else {
                m->stage.resize((size_t)n * NFIELDS);
                // rows not selected by the mask must keep the host's values: start from them
                for (int64_t l = 0; l < n; ++l)
                    std::memcpy(&m->stage[(size_t)l * NFIELDS], particles + m->gid[k][(size_t)l] * ld, sizeof(double) * NFIELDS);
                M_ENG(m, k, vpmb200_download(m->eng[k], m->stage.data(), NFIELDS, n, mask));
                for (int64_t l = 0; l < n; ++l)
                    std::memcpy(particles + m->gid[k][(size_t)l] * ld, &m->stage[(size_t)l * NFIELDS], sizeof(double) * NFIELDS);
            }
        }
        off += n;
    }
    return VPMB200_OK;
}

}  // namespace

extern "C" {

static int32_t multi_add_impl(vpmb200_multi* m, const double* cols, int64_t ld, int64_t n);
static int32_t multi_remove_impl(vpmb200_multi* m, int64_t i);
static int32_t multi_drop_statics(vpmb200_multi* m);

const char* vpmb200_multi_last_error(vpmb200_multi_handle m) { return m ? m->err.c_str() : g_multi_create_error.c_str(); }

int32_t vpmb200_multi_destroy(vpmb200_multi_handle m) {
    if (!m) return VPMB200_EINVAL;
    for (int k = 0; k < (int)m->eng.size(); ++k) {
        cudaSetDevice(m->dev[k]);
        if (m->eng[k] && m->eng[k]->stream) cudaStreamSynchronize(m->eng[k]->stream);
    }
    for (int k = 0; k < (int)m->eng.size(); ++k) {
        cudaSetDevice(m->dev[k]);
        for (auto* v : {&m->lb_send, &m->lb_recv, &m->lb_cells, &m->lb_M, &m->lb_rec, &m->lb_out, &m->lb_res})
            if (k < (int)v->size() && (*v)[k].p) cudaFree((*v)[k].p);
        if (k < (int)m->tiles_own.size()) cudaFree(m->tiles_own[k]);
        if (k < (int)m->tiles_peer.size()) cudaFree(m->tiles_peer[k]);
        if (k < (int)m->ev_pack.size() && m->ev_pack[k]) cudaEventDestroy(m->ev_pack[k]);
        if (k < (int)m->ev_done.size() && m->ev_done[k]) cudaEventDestroy(m->ev_done[k]);
        if (m->eng[k]) vpmb200_destroy(m->eng[k]);
    }
    delete m;
    return VPMB200_OK;
}

int32_t vpmb200_multi_create(int64_t max_particles, int32_t nfields, int32_t float_bits, int32_t ngpus, const int32_t* devices,
                             vpmb200_multi_handle* out) {
    if (!out) return VPMB200_EINVAL;
    *out = nullptr;
    if (ngpus < 1 || ngpus > 64) return mfail(nullptr, VPMB200_EINVAL, "ngpus must be in 1..64");
    if (max_particles <= 0) return mfail(nullptr, VPMB200_EINVAL, "max_particles must be positive");
    vpmb200_multi* m = new (std::nothrow) vpmb200_multi();
    if (!m) return VPMB200_EINVAL;
    m->G = ngpus;
    m->maxp = max_particles;
    // a shard may hold more than its even share (particles are appended to the least-loaded shard, removed anywhere)
    m->cap_local = std::min<int64_t>(max_particles, (max_particles + ngpus - 1) / ngpus * 5 / 4 + 1024);
    m->slot_tiles = vpmb200_tiles_for(m->cap_local);
    vpmb200_default_schemes(&m->sch);
    m->gid.assign(ngpus, {});
    for (int k = 0; k < ngpus; ++k) {
        m->dev.push_back(devices ? devices[k] : k);
        vpmb200_engine* e = nullptr;
        int32_t rc = vpmb200_create(m->cap_local, nfields, float_bits, m->dev[k], &e);
        m->eng.push_back(e);
        if (rc) {
            g_multi_create_error = "GPU " + std::to_string(m->dev[k]) + ": " + vpmb200_last_error(nullptr);
            vpmb200_multi_destroy(m);
            return rc;
        }
    }
    for (int k = 0; k < ngpus; ++k) {
        cudaSetDevice(m->dev[k]);
        for (int q = 0; q < ngpus; ++q) {   // direct peer copies over NVLink where the topology allows (else the runtime stages)
            if (m->dev[q] == m->dev[k]) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, m->dev[k], m->dev[q]) == cudaSuccess && can) {
                cudaError_t st = cudaDeviceEnablePeerAccess(m->dev[q], 0);
                if (st != cudaSuccess) cudaGetLastError();   // already enabled is fine
            }
        }
        double *a = nullptr, *b = nullptr;
        cudaEvent_t e1 = nullptr, e2 = nullptr;
        const size_t own = sizeof(double) * (size_t)m->slot_tiles * TILE_DOUBLES;
        if (cudaMalloc(&a, own) != cudaSuccess || cudaMalloc(&b, own * (size_t)ngpus) != cudaSuccess ||
            cudaEventCreateWithFlags(&e1, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&e2, cudaEventDisableTiming) != cudaSuccess) {
            g_multi_create_error = std::string("GPU ") + std::to_string(m->dev[k]) + ": " + cudaGetErrorString(cudaGetLastError());
            m->tiles_own.push_back(a); m->tiles_peer.push_back(b); m->ev_pack.push_back(e1); m->ev_done.push_back(e2);
            vpmb200_multi_destroy(m);
            return VPMB200_ECUDA;
        }
        m->tiles_own.push_back(a);
        m->tiles_peer.push_back(b);
        m->ev_pack.push_back(e1);
        m->ev_done.push_back(e2);
    }
    *out = m;
    return VPMB200_OK;
}

int32_t vpmb200_multi_set_schemes(vpmb200_multi_handle m, const vpmb200_schemes* s) {
    if (!m) return VPMB200_EINVAL;
    if (!s) return mfail(m, VPMB200_EINVAL, "schemes is NULL");
    if (s->viscous == VPMB200_VISCOUS_CORESPREADING && s->cs_sgm0 > 0)
        return mfail(m, VPMB200_ENOTSUP, "CoreSpreading's RBF re-fit is single-GPU; use cs_sgm0 = 0 on a multi-GPU handle");
    for (int k = 0; k < m->G; ++k) M_ENG(m, k, vpmb200_set_schemes(m->eng[k], s));
    m->sch = *s;
    return VPMB200_OK;
}

int32_t vpmb200_multi_set_time(vpmb200_multi_handle m, double t, int64_t nt) {
    if (!m) return VPMB200_EINVAL;
    m->t = t;
    m->nt = nt;
    for (int k = 0; k < m->G; ++k) vpmb200_set_time(m->eng[k], t, nt);
    return VPMB200_OK;
}

int32_t vpmb200_multi_get_time(vpmb200_multi_handle m, double* t, int64_t* nt) {
    if (!m) return VPMB200_EINVAL;
    if (t) *t = m->t;
    if (nt) *nt = m->nt;
    return VPMB200_OK;
}

int32_t vpmb200_multi_get_np(vpmb200_multi_handle m, int64_t* np) {
    if (!m || !np) return VPMB200_EINVAL;
    *np = m->np - m->nstatic;          // a parked static set is not part of the host's field
    return VPMB200_OK;
}

int32_t vpmb200_multi_shard_sizes(vpmb200_multi_handle m, int64_t* n_per_gpu) {
    if (!m || !n_per_gpu) return VPMB200_EINVAL;
    for (int k = 0; k < m->G; ++k) n_per_gpu[k] = (int64_t)m->gid[k].size();   // (includes a parked static set)
    return VPMB200_OK;
}

// Replace the field by columns 0..np-1: block partition (the host's order = shard 0's particles, then shard 1's, ...).
// With np equal to the current count and a partial mask, only those groups are refreshed and the sharding is kept.
int32_t vpmb200_multi_upload(vpmb200_multi_handle m, const double* particles, int64_t ld, int64_t np, uint32_t field_mask) {
    if (!m) return VPMB200_EINVAL;
    if (np < 0 || (np > 0 && !particles) || ld < NFIELDS) return mfail(m, VPMB200_EINVAL, "upload: bad arguments");
    { int32_t rcs = multi_drop_statics(m); if (rcs) return rcs; }
    if (np > m->maxp) return mfail(m, VPMB200_ECAPACITY, "np exceeds max_particles");
    if (np != m->np || (field_mask & VPMB200_FM_ALL) == VPMB200_FM_ALL) {
        int64_t lo = 0;
        for (int k = 0; k < m->G; ++k) {
            const int64_t n = np / m->G + (k < np % m->G ? 1 : 0);
            m->gid[k].resize((size_t)n);
            for (int64_t l = 0; l < n; ++l) m->gid[k][(size_t)l] = lo + l;
            lo += n;
        }
        m->np = np;
        multi_rebuild_inverse(m);
    }
    return multi_transfer(m, const_cast<double*>(particles), ld, field_mask, true);
}

int32_t vpmb200_multi_download(vpmb200_multi_handle m, double* particles, int64_t ld, int64_t np, uint32_t field_mask) {
    if (!m) return VPMB200_EINVAL;
    { int32_t rcs = multi_drop_statics(m); if (rcs) return rcs; }
    if (np != m->np) return mfail(m, VPMB200_EINVAL, "download: np must equal the field's particle count on a multi-GPU handle");
    if (np == 0) return VPMB200_OK;
    if (!particles || ld < NFIELDS) return mfail(m, VPMB200_EINVAL, "download: bad arguments");
    return multi_transfer(m, particles, ld, field_mask, false);
}

// vpm.add_particle x n: the new particles take the global indices np, np + 1, ... and live on the least-loaded shard
int32_t vpmb200_multi_add_particles(vpmb200_multi_handle m, const double* cols, int64_t ld, int64_t n) {
    if (!m) return VPMB200_EINVAL;
    if (n < 0 || (n > 0 && !cols)) return mfail(m, VPMB200_EINVAL, "add_particles: bad arguments");
    int32_t rcs = multi_drop_statics(m);          // any mutation invalidates a parked static set (as on the single handle)
    if (rcs) return rcs;
    return multi_add_impl(m, cols, ld, n);
}

static int32_t multi_add_impl(vpmb200_multi* m, const double* cols, int64_t ld, int64_t n) {
    if (m->np + n > m->maxp) return mfail(m, VPMB200_ECAPACITY, "adding particles would exceed max_particles");
    int64_t done = 0;
    while (done < n) {
        int k = 0;
        for (int q = 1; q < m->G; ++q)
            if (m->gid[q].size() < m->gid[k].size()) k = q;
        const int64_t room = m->cap_local - (int64_t)m->gid[k].size();
        if (room <= 0) return mfail(m, VPMB200_ECAPACITY, "every shard is full: call vpmb200_multi_rebalance or raise max_particles");
        const int64_t take = std::min(room, n - done);
        M_ENG(m, k, vpmb200_add_particles(m->eng[k], cols + done * ld, ld, take));
        for (int64_t j = 0; j < take; ++j) {
            m->owner.push_back(k);
            m->local.push_back((int64_t)m->gid[k].size());
            m->gid[k].push_back(m->np + done + j);
        }
        done += take;
    }
    m->np += n;
    return VPMB200_OK;
}

// vpm.remove_particle(pfield, i) on the GLOBAL order: the particle with the last global index takes index i
int32_t vpmb200_multi_remove_particle(vpmb200_multi_handle m, int64_t i) {
    if (!m) return VPMB200_EINVAL;
    int32_t rcs = multi_drop_statics(m);
    if (rcs) return rcs;
    return multi_remove_impl(m, i);
}

static int32_t multi_remove_impl(vpmb200_multi* m, int64_t i) {
    if (i < 0 || i >= m->np) return mfail(m, VPMB200_EINVAL, "particle index out of range");
    const int ka = m->owner[(size_t)i];
    const int64_t la = m->local[(size_t)i];
    M_ENG(m, ka, vpmb200_remove_particle(m->eng[ka], la));          // shard-local swap-remove: its last slot moves into la
    auto& g = m->gid[ka];
    const int64_t moved = g.back();
    g[(size_t)la] = moved;
    g.pop_back();
    if (moved != i) m->local[(size_t)moved] = la;
    const int64_t last = m->np - 1;
    if (i != last) {                                                 // rename global index `last` -> i
        const int kb = m->owner[(size_t)last];
        const int64_t lb = m->local[(size_t)last];
        m->gid[kb][(size_t)lb] = i;
        m->owner[(size_t)i] = kb;
        m->local[(size_t)i] = lb;
    }
    m->owner.pop_back();
    m->local.pop_back();
    m->np -= 1;
    return VPMB200_OK;
}

// Wake treatments (vpmb200_remove_where) over the shards: every shard compacts itself on its device; the global order the
// reference's loop would leave behind is applied to the index map on the host.
int32_t vpmb200_multi_remove_where(vpmb200_multi_handle m, int32_t criterion, const double* params, int64_t* removed) {
    if (!m) return VPMB200_EINVAL;
    if (!params || criterion < 1 || criterion > 4) return mfail(m, VPMB200_EINVAL, "remove_where: bad arguments");
    { int32_t rcs = multi_drop_statics(m); if (rcs) return rcs; }
    if (removed) *removed = 0;
    if (m->np == 0) return VPMB200_OK;
    static const int nparams[5] = {0, 2, 2, 9, 4};
    RemoveCriterion c;
    c.kind = criterion;
    for (int k = 0; k < 9; ++k) c.p[k] = k < nparams[criterion] ? params[k] : 0.0;
    std::vector<char> keep_global((size_t)m->np, 1);
    std::vector<std::vector<char>> keep_local(m->G);
    for (int k = 0; k < m->G; ++k) {
        vpmb200_engine* e = m->eng[k];
        const int64_t n = e->np;
        keep_local[k].assign((size_t)n, 1);
        if (n == 0) continue;
        M_CU(m, cudaSetDevice(e->device));
        std::string err;
        if (fmm_reserve_particles(e->fmm, n, err) != cudaSuccess) return mfail(m, VPMB200_ECUDA, err);
        e->shard_sorted_np = -1;
        keep_flags_kernel<<<blocks_for(n, PK_BT), PK_BT, 0, e->stream>>>(e->state, e->ld, n, c, e->fmm.perm);
        M_CU(m, cudaGetLastError());
        e->launches++;
        std::vector<int> h((size_t)n);
        M_CU(m, cudaMemcpyAsync(h.data(), e->fmm.perm, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, e->stream));
        M_CU(m, cudaStreamSynchronize(e->stream));
        for (int64_t l = 0; l < n; ++l) {
            keep_local[k][(size_t)l] = (char)h[(size_t)l];
            keep_global[(size_t)m->gid[k][(size_t)l]] = (char)h[(size_t)l];
        }
    }
    const std::vector<int64_t> a = reference_loop_order(keep_global);   // a[new global index] = old global index
    const int64_t K = (int64_t)a.size();
    if (K == m->np) return VPMB200_OK;
    std::vector<int64_t> newpos((size_t)m->np, -1);
    for (int64_t p = 0; p < K; ++p) newpos[(size_t)a[(size_t)p]] = p;
    for (int k = 0; k < m->G; ++k) {
        const std::vector<int64_t> al = reference_loop_order(keep_local[k]);   // the order vpmb200_remove_where leaves on the shard
        int64_t rm = 0;
        M_ENG(m, k, vpmb200_remove_where(m->eng[k], criterion, params, &rm));
        if ((int64_t)al.size() != m->eng[k]->np) return mfail(m, VPMB200_ECUDA, "remove_where: shard count disagrees with the host replay");
        std::vector<int64_t> g(al.size());
        for (size_t p = 0; p < al.size(); ++p) g[p] = newpos[(size_t)m->gid[k][(size_t)al[p]]];
        m->gid[k].swap(g);
    }
    if (removed) *removed = m->np - K;
    m->np = K;
    multi_rebuild_inverse(m);
    return VPMB200_OK;
}

// Move the tail of the fullest shard to the emptiest one (device to device) until max - min <= max(tolerance * mean, 1).
// The global order does not change.  *moved: particles moved.
int32_t vpmb200_multi_rebalance(vpmb200_multi_handle m, double tolerance, int64_t* moved_out) {
    if (!m) return VPMB200_EINVAL;
    if (moved_out) *moved_out = 0;
    { int32_t rcs = multi_drop_statics(m); if (rcs) return rcs; }
    if (m->G < 2) return VPMB200_OK;
    int64_t moved_total = 0;
    for (int iter = 0; iter < 4 * m->G; ++iter) {
        int hi = 0, lo = 0;
        for (int k = 1; k < m->G; ++k) {
            if (m->gid[k].size() > m->gid[hi].size()) hi = k;
            if (m->gid[k].size() < m->gid[lo].size()) lo = k;
        }
        const int64_t nh = (int64_t)m->gid[hi].size(), nl = (int64_t)m->gid[lo].size();
        const double mean = (double)m->np / m->G;
        if (nh - nl <= std::max<int64_t>((int64_t)(tolerance * mean), 1)) break;
        const int64_t mv = std::min<int64_t>((nh - nl) / 2, m->cap_local - nl);
        if (mv <= 0) break;
        vpmb200_engine *src = m->eng[hi], *dst = m->eng[lo];
        // order: dst's stream first finishes what it has queued, then the copies run on src's stream, then both wait
        M_CU(m, cudaSetDevice(src->device));
        M_CU(m, cudaStreamSynchronize(src->stream));
        M_CU(m, cudaSetDevice(dst->device));
        M_CU(m, cudaStreamSynchronize(dst->stream));
        for (int f = 0; f < NFIELDS; ++f)
            M_CU(m, cudaMemcpyPeerAsync(dst->state + (size_t)f * dst->ld + nl, dst->device, src->state + (size_t)f * src->ld + (nh - mv),
                                        src->device, sizeof(double) * (size_t)mv, dst->stream));
        M_CU(m, cudaStreamSynchronize(dst->stream));
        src->np -= mv;
        dst->np += mv;
        src->shard_sorted_np = dst->shard_sorted_np = -1;
        drop_statics(src);
        drop_statics(dst);
        for (int64_t j = 0; j < mv; ++j) {
            const int64_t g = m->gid[hi][(size_t)(nh - mv + j)];
            m->owner[(size_t)g] = lo;
            m->local[(size_t)g] = nl + j;
            m->gid[lo].push_back(g);
        }
        m->gid[hi].resize((size_t)(nh - mv));
        moved_total += mv;
    }
    if (moved_out) *moved_out = moved_total;
    return VPMB200_OK;
}

int32_t vpmb200_multi_uj(vpmb200_multi_handle m, int32_t reset, int32_t reset_sfs, int32_t sfs) {
    if (!m) return VPMB200_EINVAL;
    if (m->nstatic > 0 && m->static_gen != m->nt) { int32_t rcs = multi_drop_statics(m); if (rcs) return rcs; }
    return multi_uj(m, reset, reset_sfs, sfs);
}

int32_t vpmb200_multi_sfs(vpmb200_multi_handle m, double a, double b) {
    if (!m) return VPMB200_EINVAL;
    return multi_sfs(m, a, b);
}

// drop the parked statics: they are the last `nstatic` global indices (nothing may be appended behind them while they are parked)
static int32_t multi_drop_statics(vpmb200_multi* m) {
    while (m->nstatic > 0) {
        int32_t rc = multi_remove_impl(m, m->np - 1);
        if (rc) return rc;
        m->nstatic--;
    }
    m->static_gen = -1;
    return VPMB200_OK;
}

// vpmb200_set_statics on a sharded field (simulation.jl:355-365): the columns join the field (least-loaded shard, static flag
// forced) for the step whose counter equals `generation`; vpmb200_multi_nextstep removes them again.
int32_t vpmb200_multi_set_statics(vpmb200_multi_handle m, const double* cols, int64_t ld, int64_t n, int64_t generation) {
    if (!m) return VPMB200_EINVAL;
    if (n < 0 || (n > 0 && !cols) || ld < NFIELDS) return mfail(m, VPMB200_EINVAL, "set_statics: bad arguments");
    int32_t rc = multi_drop_statics(m);
    if (rc) return rc;
    if (n == 0 || generation != m->nt) return VPMB200_OK;       // a set for another step is never used
    std::vector<double> tmp((size_t)n * NFIELDS);
    for (int64_t i = 0; i < n; ++i) {
        std::memcpy(&tmp[(size_t)i * NFIELDS], cols + i * ld, sizeof(double) * NFIELDS);
        tmp[(size_t)i * NFIELDS + F_STATIC] = 1.0;
    }
    if ((rc = multi_add_impl(m, tmp.data(), NFIELDS, n))) return rc;
    m->nstatic = n;
    m->static_gen = generation;
    return VPMB200_OK;
}

int32_t vpmb200_multi_nextstep(vpmb200_multi_handle m, double dt, const double* Uinf, int32_t relax) {
    if (!m) return VPMB200_EINVAL;
    if (!Uinf) return mfail(m, VPMB200_EINVAL, "Uinf is NULL");
    int32_t rc;
    if (m->nstatic > 0 && m->static_gen != m->nt && (rc = multi_drop_statics(m))) return rc;   // stale set
    if (m->np > 0) {
        if (m->sch.integration == VPMB200_INTEGRATION_EULER) {
            if ((rc = multi_sfs(m, 1.0, 1.0))) return rc;
            if ((rc = multi_stage(m, relax ? VPMB200_STAGE_UPDATE_EULER_RELAX : VPMB200_STAGE_UPDATE, 0.0, 1.0, dt, Uinf))) return rc;
        } else {
            static const double AB[3][2] = {{0.0, 1.0 / 3.0}, {-5.0 / 9.0, 15.0 / 16.0}, {-153.0 / 128.0, 8.0 / 15.0}};
            if ((rc = multi_stage(m, VPMB200_STAGE_ZERO_M))) return rc;
            for (int st = 0; st < 3; ++st) {
                if ((rc = multi_sfs(m, AB[st][0], AB[st][1]))) return rc;
                if ((rc = multi_stage(m, VPMB200_STAGE_UPDATE, AB[st][0], AB[st][1], dt, Uinf))) return rc;
            }
            if (relax && m->sch.relaxation != VPMB200_RELAX_NONE) {
                if ((rc = multi_uj(m, 1, 0, 0))) return rc;
                if ((rc = multi_stage(m, VPMB200_STAGE_RELAX))) return rc;
            }
        }
    }
    if ((rc = multi_drop_statics(m))) return rc;   // the step consumed them (the reference removes them after nextstep)
    m->t += dt;
    m->nt += 1;
    for (int k = 0; k < m->G; ++k) vpmb200_set_time(m->eng[k], m->t, m->nt);
    return VPMB200_OK;
}

// Velocity (and J) at m probe points from every shard's particles: each device evaluates its own sources, the host adds the
// shards' contributions in shard order.
int32_t vpmb200_multi_uj_probe(vpmb200_multi_handle m, const double* X, int64_t np_probe, double* U, double* J) {
    if (!m) return VPMB200_EINVAL;
    if (np_probe < 0 || (np_probe > 0 && (!X || !U))) return mfail(m, VPMB200_EINVAL, "uj_probe: bad arguments");
    if (np_probe == 0) return VPMB200_OK;
    if (m->nstatic > 0 && m->static_gen != m->nt) { int32_t rcs = multi_drop_statics(m); if (rcs) return rcs; }
    std::vector<double> u((size_t)np_probe * 3), j(J ? (size_t)np_probe * 9 : 0);
    std::fill(U, U + np_probe * 3, 0.0);
    if (J) std::fill(J, J + np_probe * 9, 0.0);
    for (int k = 0; k < m->G; ++k) {
        if (m->eng[k]->np == 0) continue;
        M_ENG(m, k, vpmb200_uj_probe(m->eng[k], X, np_probe, u.data(), J ? j.data() : nullptr));
        for (int64_t q = 0; q < np_probe * 3; ++q) U[q] += u[(size_t)q];
        if (J)
            for (int64_t q = 0; q < np_probe * 9; ++q) J[q] += j[(size_t)q];
    }
    return VPMB200_OK;
}

int32_t vpmb200_multi_synchronize(vpmb200_multi_handle m) {
    if (!m) return VPMB200_EINVAL;
    for (int k = 0; k < m->G; ++k) M_ENG(m, k, vpmb200_synchronize(m->eng[k]));
    return VPMB200_OK;
}

int32_t vpmb200_multi_engine(vpmb200_multi_handle m, int32_t k, vpmb200_handle* out) {
    if (!m || !out || k < 0 || k >= m->G) return VPMB200_EINVAL;
    *out = m->eng[k];
    return VPMB200_OK;
}

}  // extern "C"
