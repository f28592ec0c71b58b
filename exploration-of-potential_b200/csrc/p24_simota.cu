// p24_simota.cu — the fused YOLOX-24p SimOTA assignment + loss-sum path for sm_100a.
//
// Replaces, for a whole batch and without ever writing the G x A cost matrix to memory:
//   Loss_Function.get_assignments / get_in_boxes_info / pts_in_poly   models/losses.py:359-592
//   utils.boxes.bboxes_iou + pairwise circle_inter                   utils/boxes.py:102-243
//   Loss_Function.dynamic_k_matching                                 models/losses.py:444-494
//   the loss sums and re-weighting of Loss_Function.forward          models/losses.py:246-345
//
// Kernel chain (caller's stream, no host synchronisation, programmatic dependent launch between the stages; the device
// code is p24_simota_kernels.inc, compiled twice: decoded rows / raw per-level head planes):
//   k_prep   two GTs per CTA: nlabel (losses.py:190), the per-GT records (ray lengths, inscribed / reject radii), the
//            GT's centre-window pairs (appended to the batch's list), a per-image barrier, then the seeds: a handful of
//            anchors that are certainly candidates (polygon tips, centre windows and inscribed discs of the farthest
//            other GTs) are evaluated; the 10th best value T is a certified lower bound of the GT's 10th largest pair
//            value, and far2 the squared centre distance below which no prediction whatsoever can reach T (the bound H*
//            depends on the GT and the distance only).  What it writes is double-buffered by step parity: with
//            P24_F_EARLY_PREP it works beside the previous step's k_tail.
//   k_pass   persistent CTAs drawing work items from ticket counters:
//            - anchor tiles (256 anchors): THE pass over the head output (cp.async reads of the 27 geometry channels
//              of every row, or coalesced planar loads + decode).  Candidate mask (polygon test OR centre window) with
//              geometric pruning and an atan2-free angle test; for the few (GT, candidate) pairs beyond far2 a cheap
//              closed-form upper bound of the pair value, appended to the GT's top-10 list when it reaches T (overflow
//              of a tile's own work list goes to a batch-wide far queue); sum BCEWithLogits(obj, 0); outputs
//              initialised to background;
//            - window chunks (32 centre-window pairs): polygon test, exact pair value and SimOTA cost -> the GT's
//              window cost table;
//            - far-queue chunks (256 queued pairs), once every tile is done
//   k_tail   a cluster of 8 CTAs per image: dynamic k = clamp(int(sum of the 10 largest pair values), 1) per GT from
//            the bracket of the list's bounds (exact evaluation of the survivors only when the floors differ), the k
//            smallest costs of the window table -> claims, conflict resolution (argmin over all GTs), fg_mask /
//            matched_gt / pred_iou, the loss terms of the foreground anchors; the batch's last CTA reduces the sums and
//            applies the normalisation and the stateful re-weighting (losses.py:280-345), or publishes the sums to the
//            peers' mailboxes (several GPUs)
//   k_fin    (several GPUs) one warp on a side stream: waits for all ranks' sums in its mailbox, adds them in rank
//            order, finalizes.  It overlaps the next step's k_prep / k_pass, which do not depend on the global sums.
//
// Compile with -fmad=false: the discrete decisions hang on fp32 thresholds evaluated in the
// reference's operation order; bounds and fast paths use explicit fmaf.
#include "p24_common.cuh"
#include "p24_host.h"
#include <string.h>

#include <mutex>
#include <unordered_map>

namespace {

struct Level {
    int off, W, H;  // anchors [off, off + W * H) form a W x H grid, row-major (yolo_head_24p.py:222-230)
    float st;       // the level's stride (expanded_strides of its anchors)
};

struct Params {
    const float* outputs;            // decoded head output [B, A, 27 + nc] (rows), or NULL:
    long long img_stride, row_stride;
    const float* raw[3][P24_MAX_LEVELS];   // raw conv outputs per level (reg [B,26,H,W], obj [B,1,H,W], cls [B,nc,H,W]) ...
    long long raw_bs[3][P24_MAX_LEVELS];   // ... and their batch strides; decoded on load (yolo_head_24p.py:212-237)
    int B, A, nc;
    const float* labels;
    long long lab_img_stride, lab_row_stride;
    int Lmax;
    const float* x_shifts;
    const float* y_shifts;
    const float* strides;
    uint8_t* fg_mask;
    int32_t* matched_gt;
    float* pred_iou;
    int32_t* num_fg;
    int32_t* num_gt;
    int32_t* dyn_k;
    float* sums28;
    float* state26;  // optional fused finalize
    float* result54;
    float* weights27;
    // workspace
    unsigned* ticket;
    long long* acc_fix;
    int* status;
    int* seed_done;
    int* ncand;
    int* lcount;
    float* gt_rec;
    float* wtab;
    float2* list;
    unsigned* cbits;
    int2* wlist;
    int* ngt;      // num_gt per image (workspace copy of this step's parity)
    int tk_wtot;   // ticket word of this step's centre-window pair count
    int tk_item;   // ticket word of this step's tile queue
    int early;     // k_prep works before it waits for the previous grid
    int2* fq;      // the batch's far queue (GT slot, anchor): overflow of the tiles' own far-pair lists
    int fq_cap;
    float* brute;
    int* claimg;
    int* kreq;
    int* ntake;
    int* rare;
    unsigned flags;
    float* mbox[P24_MAX_RANKS];  // peer mailboxes of the fused all-reduce (nranks > 1)
    int rank, nranks;
    unsigned epoch;
    int tiles;
    int nlev;
    Level lev[P24_MAX_LEVELS];
};

// Debug-only phase timers (-DP24_TIMING): %globaltimer stamps at phase boundaries, read back with p24_debug_read_timers
#ifdef P24_TIMING
#define TM_ROWS 8192
#define TM_SLOTS 16
__device__ unsigned long long g_tstamp[3][TM_ROWS][TM_SLOTS];
__device__ __forceinline__ void tmark(int kern, int row, int slot) {
    if (row >= 0 && row < TM_ROWS) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_tstamp[kern][row][slot] = t;
    }
}
#define TMARK(kern, row, slot) do { if ((threadIdx.x & 31) == 0) tmark(kern, row, slot); } while (0)
#define TMARK0(kern, row, slot) do { if (threadIdx.x == 0) tmark(kern, row, slot); } while (0)
#else
#define TMARK(kern, row, slot)
#define TMARK0(kern, row, slot)
#endif

#define TK_ITEM 0   // ticket words
#define TK_WTOT1 8  // (second parity of TK_WTOT)
#define TK_ITEM1 9  // (second parity of TK_ITEM)
#define TK_TAIL 2
#define TK_FQ 3     // far queue: entries reserved
#define TK_WIN 4    // window-chunk queue head
#define TK_WTOT 5   // centre-window pairs of the batch (k_prep)
#define TK_FQH 6    // far queue: chunk head
#define TK_TDONE 7  // anchor tiles complete

#define P24_HEAD_RAW 0
namespace rows {
#include "p24_simota_kernels.inc"
}  // namespace rows
#undef P24_HEAD_RAW
#define P24_HEAD_RAW 1
namespace rawlv {
#include "p24_simota_kernels.inc"
}  // namespace rawlv
#undef P24_HEAD_RAW
using rows::finalize_warp;

// k_fin (several GPUs): collect side of the fused all-reduce (p24_comm_finish).  One warp: wait for the flag of every rank
// in my own mailbox, add the contributions in rank order (the same bits on every rank), finalize.  It spins without a
// time-out, like a NCCL kernel: a wrong loss is worse than a hang that the framework's watchdog reports.  The wait is
// measured (status word) so that rank skew can be told from link latency.  The host side launches it on a side stream
// with NO stream dependency on the chain (an event between two steps would break their programmatic overlap): it waits on
// a flag that the chain's last CTA sets.  The next step's kernels do not depend on the global sums and run meanwhile.
struct FinParams {
    float* mbox;       // my own mailbox
    int nranks;
    unsigned epoch;
    float* sums28;
    float* state26;
    float* result54;
    float* weights27;
    int* status;       // (may be NULL)
};

__global__ void __launch_bounds__(32) k_fin(const FinParams p) {
    const int lane = threadIdx.x;
    __shared__ float s_sums[28];
    const unsigned ep = p.epoch;
    const int half = (int)(ep % MBOX_EPOCHS) * P24_MAX_RANKS;
    // launched on a side stream without any stream dependency on the chain: every (value, epoch) word of every rank --
    // this rank's own included -- is polled until it carries the epoch (a one-warp kernel: it cannot keep the chain from
    // running); the values are added in rank order
    // (the wait that is reported runs from the FIRST contribution of any rank to the last one: the skew between the ranks
    // plus the link latency, not the time this early-launched kernel spent waiting for its own chain)
    if (lane == 0) {
        bool any = false;
        while (!any) {
            for (int r = 0; r < p.nranks && !any; ++r) {
                volatile unsigned long long* w = reinterpret_cast<volatile unsigned long long*>(p.mbox + (half + r) * MBOX_SLOT);
                any = (unsigned)(*w >> 32) == ep;
            }
            if (!any) __nanosleep(32);
        }
    }
    __syncwarp();
    const long long t0 = clock64();
    float t = 0.0f;
    if (lane < 28) {
        for (int r = 0; r < p.nranks; ++r) {
            volatile unsigned long long* w = reinterpret_cast<volatile unsigned long long*>(p.mbox + (half + r) * MBOX_SLOT + 2 * lane);
            unsigned long long v = *w;
            while ((unsigned)(v >> 32) != ep) {
                __nanosleep(32);
                v = *w;
            }
            t += __uint_as_float((unsigned)v);
        }
        s_sums[lane] = t;
        p.sums28[lane] = t;
    }
    __syncwarp();
    if (lane == 0 && p.status) p.status[ST_WAITCYC] = (int)min((long long)0x7fffffff, clock64() - t0);
    __syncwarp();
    if (p.state26) finalize_warp(s_sums, p.state26, p.result54, p.weights27);
    __threadfence();
    __syncwarp();
    if (lane == 0) *reinterpret_cast<volatile unsigned*>(p.mbox + MBOX_FLAGS + 1) = ep;
}

__global__ void k_finalize(const float* __restrict__ sums28, float* __restrict__ state26, float* __restrict__ result54,
                           float* __restrict__ weights_n27) {
    finalize_warp(sums28, state26, result54, weights_n27);
}

size_t pass_smem(int Lmax) { return (size_t)Lmax * GT_REC * sizeof(float) + (size_t)2 * Lmax * sizeof(int); }
size_t tail_smem(int Lmax, int nc) {
    const size_t cap = ((size_t)Lmax * P24_TOPK + TAIL_CL - 1) / TAIL_CL;
    size_t b = (size_t)Lmax * GT_REC * sizeof(float);                 // recs
    b += (size_t)Lmax * 2 * P24_MAX_LEVELS * sizeof(int);             // window origins
    b += (size_t)Lmax * P24_TOPK * sizeof(int);                       // claims
    b += ((cap + 1) & ~(size_t)1) * sizeof(int);                      // uniq
    b += cap * sizeof(unsigned long long);                            // best
    if (27 + nc <= ROW_PAD) b += cap * ROW_PAD * sizeof(float);       // staged rows
    return b;
}

template <typename K>
cudaError_t launch(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, const Params& p, int cluster = 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (cluster > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = (unsigned)cluster;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kernel, p);
}

template <typename K>
cudaError_t launch2(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, const Params& p, int arg) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, p, arg);
}

}  // namespace

// -------------------------------------------------------------------------------------------
// C ABI
// -------------------------------------------------------------------------------------------
extern "C" size_t p24_workspace_bytes(int B, int A, int Lmax) {
    if (B <= 0 || A <= 0 || Lmax <= 0) return 0;
    return p24_layout(B, A, Lmax).total;
}

extern "C" int p24_workspace_init(void* workspace, size_t workspace_bytes, void* stream) {
    if (!workspace || ((uintptr_t)workspace & 255) != 0) return P24_E_BADARG;
    return (int)cudaMemsetAsync(workspace, 0, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int p24_read_status(void* workspace, int B, int A, int Lmax, int32_t* h_status8, void* stream) {
    if (!workspace || !h_status8 || B <= 0 || A <= 0 || Lmax <= 0) return P24_E_BADARG;
    const P24Workspace L = p24_layout(B, A, Lmax);
    char* st = (char*)workspace + L.status;
    cudaError_t e = cudaMemcpyAsync(h_status8, st, ST_WORDS * sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    // the counters restart with every read; the error bits (word 0) are sticky
    e = cudaMemsetAsync(st + sizeof(int), 0, (ST_WORDS - 1) * sizeof(int), (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaStreamSynchronize((cudaStream_t)stream);
}

namespace {
// calls made on a workspace so far (host side; the kernels double-buffer by its parity)
unsigned step_parity(const void* workspace) {
    static std::mutex mu;
    static std::unordered_map<const void*, unsigned> count;
    std::lock_guard<std::mutex> lock(mu);
    return count[workspace]++;
}

int simota_impl(const float* outputs, int64_t img_stride, int64_t row_stride, const float* const* h_raw,
                const int64_t* h_raw_bs, int B, int A,
                                     int num_classes, const float* labels, int64_t lab_img_stride,
                                     int64_t lab_row_stride, int Lmax, const float* x_shifts, const float* y_shifts,
                                     const float* strides, const int32_t* h_levels, int n_levels, uint8_t* fg_mask,
                                     int32_t* matched_gt, float* pred_iou, int32_t* num_fg, int32_t* num_gt,
                                     int32_t* dyn_k, float* sums28, float* state26, float* result54,
                                     float* weights_n27, void* workspace, size_t workspace_bytes, uint32_t flags,
                                     void* const* h_mailboxes, int rank, int nranks, uint32_t epoch, void* stream) {
    if ((!outputs && !h_raw) || !labels || !x_shifts || !y_shifts || !strides || !fg_mask || !matched_gt || !pred_iou || !num_fg ||
        !num_gt || !dyn_k || !workspace || !h_levels)
        return P24_E_BADARG;
    if (h_raw && (!h_raw_bs || 27 + num_classes > ROW_PAD)) return h_raw_bs ? P24_E_UNSUPPORTED : P24_E_BADARG;
    if (B <= 0 || A <= 0 || Lmax <= 0 || num_classes <= 0 || Lmax > 65535 || B > 65535) return P24_E_BADARG;
    if (state26 && (!sums28 || !result54 || !weights_n27)) return P24_E_BADARG;
    if (n_levels <= 0) return P24_E_BADARG;
    if (n_levels > P24_MAX_LEVELS) return P24_E_UNSUPPORTED;
    const P24Workspace L = p24_layout(B, A, Lmax);
    if (workspace_bytes < L.total) return P24_E_WORKSPACE;
    if (((uintptr_t)workspace & 255) != 0) return P24_E_BADARG;
    const size_t dyn_pass = pass_smem(Lmax), dyn_tail = tail_smem(Lmax, num_classes);
    if (dyn_pass > 160 * 1024 || dyn_tail > 200 * 1024 || p24_tiles(A) > 65535) return P24_E_UNSUPPORTED;
    char* ws = (char*)workspace;
    Params p;
    p.outputs = h_raw ? nullptr : outputs; p.img_stride = img_stride; p.row_stride = row_stride;
    for (int t = 0; t < 3; ++t)
        for (int l = 0; l < P24_MAX_LEVELS; ++l) {
            p.raw[t][l] = nullptr;
            p.raw_bs[t][l] = 0;
            if (h_raw && l < n_levels && n_levels <= P24_MAX_LEVELS) {
                if (!h_raw[t * n_levels + l]) return P24_E_BADARG;
                p.raw[t][l] = h_raw[t * n_levels + l];
                p.raw_bs[t][l] = h_raw_bs[t * n_levels + l];
            }
        }
    p.B = B; p.A = A; p.nc = num_classes;
    p.labels = labels; p.lab_img_stride = lab_img_stride; p.lab_row_stride = lab_row_stride; p.Lmax = Lmax;
    p.x_shifts = x_shifts; p.y_shifts = y_shifts; p.strides = strides;
    p.fg_mask = fg_mask; p.matched_gt = matched_gt; p.pred_iou = pred_iou;
    p.num_fg = num_fg; p.num_gt = num_gt; p.dyn_k = dyn_k; p.sums28 = sums28;
    p.state26 = state26; p.result54 = result54; p.weights27 = weights_n27;
    p.ticket = (unsigned*)(ws + L.ticket);
    p.acc_fix = (long long*)(ws + L.acc_fix);
    p.status = (int*)(ws + L.status);
    // what k_prep writes exists twice: consecutive calls on a workspace alternate (the count is kept per workspace)
    const size_t BLs = (size_t)B * Lmax;
    const int par = (int)(step_parity(workspace) & 1u);
    p.seed_done = (int*)(ws + L.seed_done) + (size_t)par * B;
    p.ngt = (int*)(ws + L.ngt) + (size_t)par * B;
    p.ncand = (int*)(ws + L.ncand);
    p.lcount = (int*)(ws + L.lcount);
    p.gt_rec = (float*)(ws + L.gt_rec) + (size_t)par * BLs * GT_REC;
    p.wtab = (float*)(ws + L.wtab) + (size_t)par * BLs * P24_WT_STRIDE;
    p.list = (float2*)(ws + L.list);
    p.cbits = (unsigned*)(ws + L.cbits);
    p.wlist = (int2*)(ws + L.wlist) + (size_t)par * BLs * 25 * P24_MAX_LEVELS;
    p.tk_wtot = par ? TK_WTOT1 : TK_WTOT;
    p.tk_item = par ? TK_ITEM1 : TK_ITEM;
    p.early = 0;
    p.fq = (int2*)(ws + L.fq);
    p.fq_cap = (int)((size_t)B * Lmax * P24_FQ_PER_GT);
    p.brute = (float*)(ws + L.brute);
    p.claimg = (int*)(ws + L.claimg);
    p.kreq = (int*)(ws + L.kreq);
    p.ntake = (int*)(ws + L.ntake);
    p.rare = (int*)(ws + L.rare);
    p.flags = flags;
    p.rank = 0; p.nranks = 1; p.epoch = 0;
    for (int r = 0; r < P24_MAX_RANKS; ++r) p.mbox[r] = nullptr;
    if (h_mailboxes && nranks > 1) {
        if (nranks > P24_MAX_RANKS || rank < 0 || rank >= nranks || !sums28) return P24_E_BADARG;
        for (int r = 0; r < nranks; ++r) {
            if (!h_mailboxes[r]) return P24_E_BADARG;
            p.mbox[r] = (float*)h_mailboxes[r];
        }
        p.rank = rank; p.nranks = nranks; p.epoch = epoch;
    }
    p.tiles = p24_tiles(A);
    p.nlev = n_levels;
    {
        // the levels must tile [0, A) in order: anchors [off, off + W * H) of level l form a W x H grid
        long long next = 0;
        for (int l = 0; l < P24_MAX_LEVELS; ++l) {
            p.lev[l].off = 0; p.lev[l].W = 1; p.lev[l].H = 0; p.lev[l].st = 1.0f;
            if (l < n_levels) {
                const int32_t* d = h_levels + 4 * l;
                float stv;
                memcpy(&stv, d + 3, sizeof(float));
                if (d[0] != next || d[1] <= 0 || d[2] <= 0 || !(stv > 0.0f)) return P24_E_BADARG;
                p.lev[l].off = d[0]; p.lev[l].W = d[1]; p.lev[l].H = d[2]; p.lev[l].st = stv;
                next += (long long)d[1] * d[2];
            }
        }
        if (next != A) return P24_E_BADARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    // two instantiations of the chain: decoded rows / raw per-level planes (p24_simota_kernels.inc)
    const bool raw = p.outputs == nullptr;
    void (*const k_prep)(const Params, int) = raw ? rawlv::k_prep : rows::k_prep;
    void (*const k_pass)(const Params) = raw ? rawlv::k_pass : rows::k_pass;
    void (*const k_tail)(const Params) = raw ? rawlv::k_tail : rows::k_tail;
    if (p24::dev_once(raw ? 1u << 1 : 1u << 0)) {  // per device: a process may drive several GPUs
        cudaFuncSetAttribute(k_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(k_prep, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(k_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    }
    const bool pdl = !(flags & P24_F_NO_PDL) && !p24::prof_on();
    cudaError_t e = cudaSuccess;
    const int n_sm = p24::dev_info().n_sm;
    p24::prof_mark(0, st);
    {
        // k_prep waits inside for the records of all CTAs of an image: one launch when the whole grid is resident at once
        // (always, at training batch sizes), else records and seeds in two launches
        const dim3 pgrid((Lmax + PREP_GTS - 1) / PREP_GTS, B);
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_prep, PREP_THREADS, dyn_pass);
        if (e != cudaSuccess) return (int)e;
        const bool fused = (long long)pgrid.x * pgrid.y <= (long long)per_sm * n_sm;
        // early mode (opt-in, P24_F_EARLY_PREP): k_prep works beside the previous step's k_tail.  Not while the stream is
        // being captured (a replayed graph repeats one parity) and not in the two-launch fallback.
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) cap = cudaStreamCaptureStatusActive;
        p.early = (fused && pdl && cap == cudaStreamCaptureStatusNone && (flags & P24_F_EARLY_PREP)) ? 1 : 0;
        e = launch2(k_prep, pgrid, dim3(PREP_THREADS), dyn_pass, st, pdl, p, fused ? 0 : 1);
        if (e != cudaSuccess) return (int)e;
        if (!fused) {
            e = launch2(k_prep, pgrid, dim3(PREP_THREADS), dyn_pass, st, pdl, p, 2);
            if (e != cudaSuccess) return (int)e;
        }
    }
    p24::prof_mark(1, st);
    {
        // one wave of persistent CTAs drawing tiles and window chunks from ticket counters
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pass, P24_THREADS, dyn_pass);
        if (e != cudaSuccess) return (int)e;
        if (per_sm < 1) return P24_E_UNSUPPORTED;
        const long long items = (long long)B * p.tiles > (long long)B * Lmax ? (long long)B * p.tiles : (long long)B * Lmax;
        const long long cap = (long long)per_sm * n_sm;
        e = launch(k_pass, dim3((unsigned)(items < cap ? items : cap)), dim3(P24_THREADS), dyn_pass, st, pdl, p);
        if (e != cudaSuccess) return (int)e;
    }
    p24::prof_mark(2, st);
    e = launch(k_tail, dim3(B * TAIL_CL), dim3(TAIL_THREADS), dyn_tail, st, pdl, p, TAIL_CL);
    if (e != cudaSuccess) return (int)e;
    p24::prof_mark(3, st);
    return (int)cudaGetLastError();
}
}  // namespace

extern "C" int p24_simota_loss_batch(const float* outputs, int64_t img_stride, int64_t row_stride, int B, int A,
                                     int num_classes, const float* labels, int64_t lab_img_stride,
                                     int64_t lab_row_stride, int Lmax, const float* x_shifts, const float* y_shifts,
                                     const float* strides, const int32_t* h_levels, int n_levels, uint8_t* fg_mask,
                                     int32_t* matched_gt, float* pred_iou, int32_t* num_fg, int32_t* num_gt,
                                     int32_t* dyn_k, float* sums28, float* state26, float* result54,
                                     float* weights_n27, void* workspace, size_t workspace_bytes, uint32_t flags,
                                     void* const* h_mailboxes, int rank, int nranks, uint32_t epoch, void* stream) {
    if (!outputs) return P24_E_BADARG;
    return simota_impl(outputs, img_stride, row_stride, nullptr, nullptr, B, A, num_classes, labels, lab_img_stride,
                       lab_row_stride, Lmax, x_shifts, y_shifts, strides, h_levels, n_levels, fg_mask, matched_gt, pred_iou,
                       num_fg, num_gt, dyn_k, sums28, state26, result54, weights_n27, workspace, workspace_bytes, flags,
                       h_mailboxes, rank, nranks, epoch, stream);
}

extern "C" int p24_simota_loss_batch_raw(const float* const* h_raw, const int64_t* h_raw_batch_stride, int B, int A,
                                         int num_classes, const float* labels, int64_t lab_img_stride,
                                         int64_t lab_row_stride, int Lmax, const float* x_shifts, const float* y_shifts,
                                         const float* strides, const int32_t* h_levels, int n_levels, uint8_t* fg_mask,
                                         int32_t* matched_gt, float* pred_iou, int32_t* num_fg, int32_t* num_gt,
                                         int32_t* dyn_k, float* sums28, float* state26, float* result54,
                                         float* weights_n27, void* workspace, size_t workspace_bytes, uint32_t flags,
                                         void* const* h_mailboxes, int rank, int nranks, uint32_t epoch, void* stream) {
    if (!h_raw || !h_raw_batch_stride) return P24_E_BADARG;
    return simota_impl(nullptr, 0, 0, h_raw, h_raw_batch_stride, B, A, num_classes, labels, lab_img_stride, lab_row_stride,
                       Lmax, x_shifts, y_shifts, strides, h_levels, n_levels, fg_mask, matched_gt, pred_iou, num_fg, num_gt,
                       dyn_k, sums28, state26, result54, weights_n27, workspace, workspace_bytes, flags, h_mailboxes, rank,
                       nranks, epoch, stream);
}

extern "C" size_t p24_comm_mailbox_bytes(void) { return (size_t)(MBOX_FLAGS + 64) * sizeof(float); }

extern "C" int p24_comm_finish(void* d_own_mailbox, int nranks, uint32_t epoch, float* sums28, float* state26, float* result54,
                               float* weights_n27, void* workspace, int B, int A, int Lmax, void* stream) {
    if (!d_own_mailbox || !sums28 || nranks < 2 || nranks > P24_MAX_RANKS) return P24_E_BADARG;
    if (state26 && (!result54 || !weights_n27)) return P24_E_BADARG;
    FinParams f;
    f.mbox = (float*)d_own_mailbox;
    f.nranks = nranks;
    f.epoch = epoch;
    f.sums28 = sums28; f.state26 = state26; f.result54 = result54; f.weights27 = weights_n27;
    f.status = nullptr;
    if (workspace && B > 0 && A > 0 && Lmax > 0) f.status = (int*)((char*)workspace + p24_layout(B, A, Lmax).status);
    k_fin<<<1, 32, 0, (cudaStream_t)stream>>>(f);
    return (int)cudaGetLastError();
}

extern "C" int p24_comm_alloc(void** d_mailbox) {
    if (!d_mailbox) return P24_E_BADARG;
    cudaError_t e = cudaMalloc(d_mailbox, p24_comm_mailbox_bytes());
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(*d_mailbox, 0, p24_comm_mailbox_bytes());
    if (e != cudaSuccess) return (int)e;
    return (int)cudaDeviceSynchronize();
}

extern "C" int p24_comm_free(void* d_mailbox) { return (int)cudaFree(d_mailbox); }

extern "C" int p24_comm_export(void* d_mailbox, void* h_handle64) {
    if (!d_mailbox || !h_handle64) return P24_E_BADARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    return (int)cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(h_handle64), d_mailbox);
}

extern "C" int p24_comm_import(const void* h_handle64, void** d_peer_mailbox) {
    if (!h_handle64 || !d_peer_mailbox) return P24_E_BADARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, h_handle64, sizeof(h));
    return (int)cudaIpcOpenMemHandle(d_peer_mailbox, h, cudaIpcMemLazyEnablePeerAccess);
}

extern "C" int p24_comm_close(void* d_peer_mailbox) { return (int)cudaIpcCloseMemHandle(d_peer_mailbox); }

extern "C" int p24_loss_finalize(const float* sums28, float* state26, float* result54, float* weights_n27,
                                 void* stream) {
    if (!sums28 || !state26 || !result54 || !weights_n27) return P24_E_BADARG;
    k_finalize<<<1, 32, 0, (cudaStream_t)stream>>>(sums28, state26, result54, weights_n27);
    return (int)cudaGetLastError();
}

#ifdef P24_TIMING
extern "C" int p24_debug_read_timers(unsigned long long* h_out) {
    return (int)cudaMemcpyFromSymbol(h_out, g_tstamp, sizeof(unsigned long long) * 3 * TM_ROWS * TM_SLOTS);
}
#endif
