// ptb_multi.inl — multi-GPU behind the C-ABI (included at the end of ptb_engine.cu: it uses the context, the pass loop and the tile
// pack kernel defined there).
//
// The reference seam is ONE call, Raytracer::render_image_nopreviz (mainApp.cpp:38-49), so the multi-GPU render sits under one
// call too.  Pixels are independent given the per-(pixel,sample) streams and the scene is read-only: every GPU holds the whole
// scene and renders the image tiles it owns (owner = tile id % n, rows rotated, ptb_scene.h); the only exchange is the gather of
// the owned tiles (+ their ceil(2 sigma) splat apron) on rank 0: k_shard_pack -> ncclSend / ncclRecv over NVLink -> unpack-add
// (vector reds) -> k_resolve, all on the context's own stream (no second stream to order against).
// Two ways in:
//   * one process per GPU (torchrun): ptb_comm_unique_id / ptb_comm_init, then every rank calls ptb_render_sharded;
//   * one process, several GPUs: ptb_group_* owns one context and one host thread per device, builds the BVH once, uploads it to
//     every device and drives the same ptb_render_sharded on each.
// NCCL is bound at run time (dlopen): the copy already in the process (torch's) if there is one, else the system libnccl.so.2.
namespace {
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};
NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {getenv("PTB_NCCL_PATH"), "libnccl.so.2", "libnccl.so"};
        api.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);          // the copy torch (or the host program) already loaded
        for (const char* n : names) if (!api.h && n && *n) api.h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (!api.h) { api.err = std::string("NCCL not found (libnccl.so.2; set PTB_NCCL_PATH): ") + (dlerror() ? dlerror() : ""); return; }
#define PTB_NCCL_SYM(field, name) do { *(void**)(&api.field) = dlsym(api.h, name); if (!api.field) api.err = std::string("NCCL symbol missing: ") + name; } while (0)
        PTB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId"); PTB_NCCL_SYM(CommInitRank, "ncclCommInitRank"); PTB_NCCL_SYM(CommInitAll, "ncclCommInitAll");
        PTB_NCCL_SYM(CommDestroy, "ncclCommDestroy"); PTB_NCCL_SYM(Send, "ncclSend"); PTB_NCCL_SYM(Recv, "ncclRecv");
        PTB_NCCL_SYM(GroupStart, "ncclGroupStart"); PTB_NCCL_SYM(GroupEnd, "ncclGroupEnd"); PTB_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef PTB_NCCL_SYM
    });
    return &api;
}
}  // namespace

#define NK(call)                                                                                         \
    do {                                                                                                 \
        ncclResult_t r_ = (call);                                                                        \
        if (r_ != ncclSuccess) {                                                                         \
            c->err = std::string(#call) + ": " + nccl_api()->GetErrorString(r_);                         \
            return PTB_ERR_CUDA;                                                                         \
        }                                                                                                \
    } while (0)

extern "C" {

int ptb_comm_unique_id(void* out_id128) {
    if (!out_id128) return PTB_ERR_INVALID;
    NcclApi* a = nccl_api();
    if (!a->err.empty()) { g_create_err = a->err; return PTB_ERR_UNSUPPORTED; }
    static_assert(sizeof(ncclUniqueId) == PTB_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    if (a->GetUniqueId(&id) != ncclSuccess) { g_create_err = "ncclGetUniqueId failed"; return PTB_ERR_CUDA; }
    memcpy(out_id128, &id, sizeof(id));
    return PTB_OK;
}

int ptb_comm_destroy(ptb_ctx* c) {
    if (!c) return PTB_ERR_INVALID;
    if (c->comm) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        if (c->comm_owned) nccl_api()->CommDestroy((ncclComm_t)c->comm);
    }
    c->comm = nullptr; c->comm_n = 1; c->comm_rank = 0; c->comm_owned = false;
    return PTB_OK;
}

int ptb_comm_init(ptb_ctx* c, int n_ranks, int rank, const void* id128) {
    if (!c || n_ranks < 1 || rank < 0 || rank >= n_ranks || (n_ranks > 1 && !id128)) { if (c) c->err = "comm_init: bad rank / size / id"; return PTB_ERR_INVALID; }
    ptb_comm_destroy(c);
    c->comm_n = n_ranks; c->comm_rank = rank;
    if (n_ranks == 1) return PTB_OK;
    NcclApi* a = nccl_api();
    if (!a->err.empty()) { c->err = a->err; return PTB_ERR_UNSUPPORTED; }
    CK(cudaSetDevice(c->device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    NK(a->CommInitRank(&comm, n_ranks, id, rank));
    c->comm = comm; c->comm_owned = true;
    return PTB_OK;
}

// Host buffers the caller keeps across frames (Raytracer::imagedouble / sample_count / image are members, Raytracer.h:90-105) can be
// page-locked once so that the device->host copies at the end of a render run at PCIe speed without a staging copy.
int ptb_pin_host_buffer(ptb_ctx* c, void* ptr, int64_t bytes) {
    if (!c || !ptr || bytes <= 0) return PTB_ERR_INVALID;
    for (auto& e : c->pinned) if (e.first == ptr) return PTB_OK;
    CK(cudaSetDevice(c->device));
    CK(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable));
    c->pinned.emplace_back(ptr, bytes);
    return PTB_OK;
}
int ptb_unpin_host_buffer(ptb_ctx* c, void* ptr) {
    if (!c || !ptr) return PTB_ERR_INVALID;
    for (size_t i = 0; i < c->pinned.size(); i++)
        if (c->pinned[i].first == ptr) {
            cudaSetDevice(c->device);
            cudaStreamSynchronize(c->stream);
            cudaHostUnregister(ptr);
            c->pinned.erase(c->pinned.begin() + (long)i);
            return PTB_OK;
        }
    return PTB_OK;
}

int ptb_render_sharded(ptb_ctx* c, const ptb_camera* cam, const ptb_params* p, float* imagedouble, float* sample_count, uint8_t* image, ptb_stats* stats) {
    if (!c || !p) return PTB_ERR_INVALID;
    if (!c->committed) { c->err = "render before commit"; return PTB_ERR_STATE; }
    if (c->info_n_tri < 0) { const int rf = ensure_frame(c); if (rf) return rf; }     // (group members are re-posed by ptb_group_render)
    CK(cudaSetDevice(c->device));
    auto w0 = std::chrono::steady_clock::now();
    const int n = c->comm_n, rank = c->comm_rank;
    ptb_params q = *p;
    q.shard_rank = rank; q.shard_count = n;
    FrameDev f;
    int rc = frame_setup(c, cam, &q, f);
    if (rc) return rc;
    const int64_t npix = (int64_t)q.W * q.H;
    if ((rc = grow(c, &c->d_accum, &c->accum_n, npix))) return rc;
    CK(cudaMemsetAsync(c->d_accum, 0, (size_t)npix * sizeof(F4), c->stream));
    if ((rc = render_passes(c, f, q.nrays, c->d_accum, stats))) return rc;
    if (n > 1) {
        NcclApi* a = nccl_api();
        int tile, apron, tiles_x, total, mine;
        std::vector<int64_t> cnt(n), off(n, 0);     // packed float4 texels per rank; offsets into rank 0's receive buffer
        int64_t recv_total = 0;
        for (int r = 0; r < n; r++) {
            shard_geometry(&q, r, tile, apron, tiles_x, total, mine);
            const int64_t side = tile + 2 * apron;
            cnt[r] = (int64_t)mine * side * side;
            if (r > 0) { off[r] = recv_total; recv_total += cnt[r]; }
        }
        shard_geometry(&q, rank, tile, apron, tiles_x, total, mine);
        const int shift = shard_tile_shift(tiles_x, n);
        if (rank != 0) {
            if (cnt[rank] > 0) {
                if ((rc = grow(c, &c->d_pack, &c->pack_n, cnt[rank]))) return rc;
                k_shard_pack<<<(unsigned)((cnt[rank] + 255) / 256), 256, 0, c->stream>>>(c->d_accum, c->d_pack, q.W, q.H, tile, apron, tiles_x, total, shift, rank, n, cnt[rank], 0);
                CK(cudaGetLastError());
                NK(a->Send(c->d_pack, (size_t)cnt[rank] * 4, ncclFloat, 0, (ncclComm_t)c->comm, c->stream));
            }
            CK(cudaStreamSynchronize(c->stream));
        } else {
            if ((rc = grow(c, &c->d_pack, &c->pack_n, std::max<int64_t>(recv_total, 1)))) return rc;
            NK(a->GroupStart());
            for (int r = 1; r < n; r++)
                if (cnt[r] > 0) NK(a->Recv(c->d_pack + off[r], (size_t)cnt[r] * 4, ncclFloat, r, (ncclComm_t)c->comm, c->stream));
            NK(a->GroupEnd());
            for (int r = 1; r < n; r++)
                if (cnt[r] > 0) {
                    k_shard_pack<<<(unsigned)((cnt[r] + 255) / 256), 256, 0, c->stream>>>(c->d_accum, c->d_pack + off[r], q.W, q.H, tile, apron, tiles_x, total, shift, r, n, cnt[r], 1);
                    CK(cudaGetLastError());
                }
        }
    }
    if (stats && n > 1) {   // device time of this rank's share INCLUDING its part of the gather (ev0: start of the pass loop)
        CK(cudaEventRecord(c->ev1, c->stream));
        CK(cudaEventSynchronize(c->ev1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        stats->ms_device = ms;
    }
    if (rank == 0) {
        if (imagedouble || sample_count || image) { if ((rc = resolve_to_host(c, c->d_accum, q.W, q.H, q.gamma, imagedouble, sample_count, image))) return rc; }
        else CK(cudaStreamSynchronize(c->stream));      // resident form: the gathered sums stay in the context (ptb_resolve_last reads them)
    }
    if (stats) stats->ms_wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
    return PTB_OK;
}

// Normalise + tonemap the sums the last ptb_render / ptb_render_sharded left on the device (rank 0) into host outputs.
int ptb_resolve_last(ptb_ctx* c, int W, int H, float gamma, float* imagedouble, float* sample_count, uint8_t* image) {
    if (!c || W <= 0 || H <= 0) return PTB_ERR_INVALID;
    if (!c->d_accum || c->accum_n < (int64_t)W * H) { c->err = "resolve_last: no frame of that size on the device"; return PTB_ERR_STATE; }
    CK(cudaSetDevice(c->device));
    return resolve_to_host(c, c->d_accum, W, H, gamma, imagedouble, sample_count, image);
}

// ---- one process, several GPUs ---------------------------------------------------------------------------------------------------
struct ptb_group {
    std::vector<ptb_ctx*> ctx;      // ctx[0] is the leader: it owns the host scene
    std::string err;
};

const char* ptb_group_last_error(const ptb_group* g) { return g ? g->err.c_str() : g_create_err.c_str(); }

void ptb_group_destroy(ptb_group* g) {
    if (!g) return;
    for (ptb_ctx* c : g->ctx) if (c) { ptb_comm_destroy(c); ptb_destroy(c); }
    delete g;
}

int ptb_group_create(const int* device_ids, int n_devices, ptb_group** out) {
    if (!out || !device_ids || n_devices < 1 || n_devices > 64) return PTB_ERR_INVALID;
    *out = nullptr;
    ptb_group* g = new ptb_group();
    for (int i = 0; i < n_devices; i++) {
        for (int j = 0; j < i; j++) if (device_ids[j] == device_ids[i]) { g_create_err = "group_create: a device is listed twice"; ptb_group_destroy(g); return PTB_ERR_INVALID; }
        ptb_ctx* c = nullptr;
        const int rc = ptb_create(device_ids[i], &c);
        if (rc) { ptb_group_destroy(g); return rc; }
        g->ctx.push_back(c);
        c->comm_n = n_devices; c->comm_rank = i;
    }
    if (n_devices > 1) {
        NcclApi* a = nccl_api();
        if (!a->err.empty()) { g_create_err = a->err; ptb_group_destroy(g); return PTB_ERR_UNSUPPORTED; }
        std::vector<ncclComm_t> comms(n_devices);
        const ncclResult_t r = a->CommInitAll(comms.data(), n_devices, device_ids);
        if (r != ncclSuccess) { g_create_err = std::string("ncclCommInitAll: ") + a->GetErrorString(r); ptb_group_destroy(g); return PTB_ERR_CUDA; }
        for (int i = 0; i < n_devices; i++) { g->ctx[i]->comm = comms[i]; g->ctx[i]->comm_owned = true; }
    }
    *out = g;
    return PTB_OK;
}

int ptb_group_size(const ptb_group* g) { return g ? (int)g->ctx.size() : 0; }
ptb_ctx* ptb_group_ctx(ptb_group* g, int i) { return (g && i >= 0 && i < (int)g->ctx.size()) ? g->ctx[i] : nullptr; }

// ptb_commit for the group: flatten + BVH build once on the leader's host scene, one upload per device (concurrently)
int ptb_group_commit(ptb_group* g) {
    if (!g || g->ctx.empty()) return PTB_ERR_INVALID;
    ptb_ctx* lead = g->ctx[0];
    int rc = commit_flatten(lead);
    if (rc) { g->err = lead->err; return rc; }
    std::vector<int> rcs(g->ctx.size(), PTB_OK);
    std::vector<std::thread> th;
    for (size_t i = 0; i < g->ctx.size(); i++)
        th.emplace_back([&, i] { rcs[i] = commit_upload(g->ctx[i], lead->flat, lead->host); });
    for (auto& t : th) t.join();
    for (size_t i = 0; i < g->ctx.size(); i++) {
        if (rcs[i]) { g->err = "device " + std::to_string(g->ctx[i]->device) + ": " + g->ctx[i]->err; return rcs[i]; }
        if (i > 0) { g->ctx[i]->info_n_tri = lead->flat.n_tri_scene; g->ctx[i]->info_nodes = lead->flat.bvh.n_nodes; g->ctx[i]->info_depth = lead->flat.bvh.depth; }
    }
    commit_release_host(lead);
    return PTB_OK;
}

// options apply to every device of the group
int ptb_group_set_option(ptb_group* g, int option, int64_t value) {
    if (!g) return PTB_ERR_INVALID;
    for (ptb_ctx* c : g->ctx) { const int rc = ptb_set_option(c, option, value); if (rc) return rc; }
    return PTB_OK;
}

// replaces: ONE Raytracer::render_image_nopreviz() call (mainApp.cpp:44) rendered by all the group's GPUs.  stats: sums over the
// devices; ms_device = the slowest device.
int ptb_group_render(ptb_group* g, const ptb_camera* cam, const ptb_params* p, float* imagedouble, float* sample_count, uint8_t* image, ptb_stats* stats) {
    if (!g || g->ctx.empty() || !cam || !p) return PTB_ERR_INVALID;
    auto w0 = std::chrono::steady_clock::now();
    const size_t n = g->ctx.size();
    std::vector<int> rcs(n, PTB_OK);
    ptb_ctx* lead = g->ctx[0];
    if (lead->frame_dirty) {     // ptb_set_frame on the leader: matrices once on the host, triangles and node boxes on every device
        refit_host(lead);
        std::vector<std::thread> rt;
        for (size_t i = 0; i < n; i++) rt.emplace_back([&, i] { rcs[i] = refit_device(g->ctx[i], lead->flat, lead->host); });
        for (auto& t : rt) t.join();
        for (size_t i = 0; i < n; i++) if (rcs[i]) { g->err = "device " + std::to_string(g->ctx[i]->device) + " (refit): " + g->ctx[i]->err; return rcs[i]; }
    }
    std::vector<ptb_stats> st(n);
    std::vector<std::thread> th;
    for (size_t i = 1; i < n; i++)
        th.emplace_back([&, i] { rcs[i] = ptb_render_sharded(g->ctx[i], cam, p, nullptr, nullptr, nullptr, &st[i]); });
    rcs[0] = ptb_render_sharded(g->ctx[0], cam, p, imagedouble, sample_count, image, &st[0]);
    for (auto& t : th) t.join();
    for (size_t i = 0; i < n; i++) if (rcs[i]) { g->err = "device " + std::to_string(g->ctx[i]->device) + ": " + g->ctx[i]->err; return rcs[i]; }
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        for (const ptb_stats& s : st) {
            stats->samples += s.samples; stats->rays_closest += s.rays_closest; stats->rays_shadow += s.rays_shadow; stats->node_visits += s.node_visits;
            stats->tri_tests += s.tri_tests; stats->kernel_launches += s.kernel_launches; stats->ms_device = std::max(stats->ms_device, s.ms_device);
        }
        stats->ms_wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
    }
    return PTB_OK;
}
}  // extern "C"
#undef NK
