// Batched non_max_suppression, bit-exact with the reference on the kept-index sets.
//
// Reference: basics/utils/general.py:425-512 and torchvision.ops.nms (general.py:496).
//
// Pipeline (all on `stream`, no host sync, no allocation):
//   K1 scan      one coalesced pass over pred [B,R,5+nc]: objectness / class-confidence thresholds,
//                class filter, candidates written (ordered inside each 256-row chunk) to a staging area
//   K2 offsets   per image: exclusive scan of the chunk counts
//   K3 compact   staging -> dense candidate list in the reference's candidate order (row-major (row, class))
//   K4 sort      per image: stable LSD radix sort by confidence, descending (shared-memory histograms)
//   K5 gather    top min(n, max_nms) candidates: xywh->xyxy, class offset (cls*max_wh), fp32 like the reference
//   K6 suppress  per image: greedy NMS in chunks of 512 with a shared-memory bitmask, early exit after
//                max_det survivors (greedy NMS is sequential in score order, so the first max_det
//                survivors do not depend on anything after them), merge-NMS, `redundant` filter, output
//
// Arithmetic that decides an index is done with explicit round-to-nearest intrinsics (no FMA
// contraction) in the reference's operation order.
#include "common.cuh"

namespace sodt {
namespace {

constexpr int CHUNK = 256;         // rows per K1 block
constexpr int SORT_THREADS = 1024;
constexpr int NMS_THREADS = 512;   // == suppression chunk
constexpr int NMS_WORDS = NMS_THREADS / 32;

struct Workspace {
    uint32_t *st_key, *st_code;    // [B][cap]  staging; reused as the sort's pong buffers
    uint32_t *cp_key, *cp_idx;     // [B][cap]  compact keys / original compact index (sort ping buffers)
    uint32_t* cp_code;             // [B][cap]  row*nc + cls per compact candidate
    int *chunk_cnt, *chunk_off;    // [B][nchunks]
    int* n_cand;                   // [B]
    float4 *box, *box_off;         // [B][max_nms]
    float* score;                  // [B][max_nms]
    int* cls;                      // [B][max_nms]
    size_t bytes;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

Workspace carve(void* base, int B, int R, int nc, int multi_label, int max_nms) {
    Workspace w;
    const size_t nchunks = (R + CHUNK - 1) / CHUNK;
    const size_t cap = nchunks * CHUNK * (size_t)(multi_label ? nc : 1);
    char* p = static_cast<char*>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off = align_up(off + bytes, 256); return r; };
    w.st_key = reinterpret_cast<uint32_t*>(take(sizeof(uint32_t) * cap * B));
    w.st_code = reinterpret_cast<uint32_t*>(take(sizeof(uint32_t) * cap * B));
    w.cp_key = reinterpret_cast<uint32_t*>(take(sizeof(uint32_t) * cap * B));
    w.cp_idx = reinterpret_cast<uint32_t*>(take(sizeof(uint32_t) * cap * B));
    w.cp_code = reinterpret_cast<uint32_t*>(take(sizeof(uint32_t) * cap * B));
    w.chunk_cnt = reinterpret_cast<int*>(take(sizeof(int) * nchunks * B));
    w.chunk_off = reinterpret_cast<int*>(take(sizeof(int) * nchunks * B));
    w.n_cand = reinterpret_cast<int*>(take(sizeof(int) * B));
    w.box = reinterpret_cast<float4*>(take(sizeof(float4) * (size_t)max_nms * B));
    w.box_off = reinterpret_cast<float4*>(take(sizeof(float4) * (size_t)max_nms * B));
    w.score = reinterpret_cast<float*>(take(sizeof(float) * (size_t)max_nms * B));
    w.cls = reinterpret_cast<int*>(take(sizeof(int) * (size_t)max_nms * B));
    w.bytes = off;
    return w;
}

__device__ __forceinline__ bool class_allowed(int j, const int* __restrict__ classes, int n_classes) {
    if (classes == nullptr) return true;
    for (int e = 0; e < n_classes; ++e)
        if (classes[e] == j) return true;
    return false;
}

// ---------------------------------------------------------------------------------------- K1
__global__ void __launch_bounds__(CHUNK)
nms_scan_kernel(const float* __restrict__ pred, const int* __restrict__ classes, int n_classes,
                uint32_t* __restrict__ st_key, uint32_t* __restrict__ st_code, int* __restrict__ chunk_cnt,
                int R, int nc, float conf_t, int multi_label, int vec_ok) {
    extern __shared__ __align__(16) float rows[];  // [CHUNK][no]
    __shared__ int warp_tot[CHUNK / 32];
    const int no = 5 + nc;
    const int chunk = blockIdx.x, b = blockIdx.y, nchunks = gridDim.x;
    const int r0 = chunk * CHUNK;
    const int nrows = min(CHUNK, R - r0);
    const float* src = pred + ((long long)b * R + r0) * no;
    const int nflt = nrows * no;
    if (vec_ok) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(rows);
        for (int e = threadIdx.x; e < nflt / 4; e += CHUNK) d4[e] = __ldg(s4 + e);
        for (int e = (nflt / 4) * 4 + threadIdx.x; e < nflt; e += CHUNK) rows[e] = __ldg(src + e);
    } else {
        for (int e = threadIdx.x; e < nflt; e += CHUNK) rows[e] = __ldg(src + e);
    }
    __syncthreads();

    const int t = threadIdx.x;
    const float* row = rows + t * no;
    int cnt = 0, best = -1;
    float best_conf = 0.f;
    const bool live = t < nrows && row[4] > conf_t;              // general.py:433
    if (live) {
        const float obj = row[4];
        if (multi_label) {
            for (int j = 0; j < nc; ++j)
                cnt += (__fmul_rn(row[5 + j], obj) > conf_t) && class_allowed(j, classes, n_classes);  // :465,:472
        } else {
            best = 0;
            best_conf = __fmul_rn(row[5], obj);
            for (int j = 1; j < nc; ++j) {
                const float c = __fmul_rn(row[5 + j], obj);
                if (c > best_conf) { best_conf = c; best = j; }   // first maximum, :475
            }
            cnt = (best_conf > conf_t) && class_allowed(best, classes, n_classes);
        }
    }
    // ordered positions inside the chunk: warp scan + warp totals
    const int lane = t & 31, wid = t >> 5;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    int base = 0;
    for (int wq = 0; wq < wid; ++wq) base += warp_tot[wq];
    int pos = base + incl - cnt;
    const int per = multi_label ? nc : 1;
    const long long slot0 = ((long long)b * nchunks + chunk) * (long long)(CHUNK * per);
    if (cnt > 0) {
        const float obj = row[4];
        const uint32_t rowcode = (uint32_t)(r0 + t) * (uint32_t)nc;
        if (multi_label) {
            for (int j = 0; j < nc; ++j) {
                const float c = __fmul_rn(row[5 + j], obj);
                if (c > conf_t && class_allowed(j, classes, n_classes)) {
                    st_key[slot0 + pos] = ~__float_as_uint(c);   // ascending radix order == descending confidence
                    st_code[slot0 + pos] = rowcode + j;
                    ++pos;
                }
            }
        } else {
            st_key[slot0 + pos] = ~__float_as_uint(best_conf);
            st_code[slot0 + pos] = rowcode + best;
        }
    }
    if (t == CHUNK - 1) chunk_cnt[b * nchunks + chunk] = base + incl;
}

// ---------------------------------------------------------------------------------------- K2
__global__ void __launch_bounds__(1024)
nms_offsets_kernel(const int* __restrict__ chunk_cnt, int* __restrict__ chunk_off, int* __restrict__ n_cand, int nchunks) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, wid = t >> 5;
    if (t == 0) carry = 0;
    __syncthreads();
    for (int c0 = 0; c0 < nchunks; c0 += 1024) {
        const int c = c0 + t;
        const int v = c < nchunks ? chunk_cnt[b * nchunks + c] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        int base = carry;
        for (int wq = 0; wq < wid; ++wq) base += warp_tot[wq];
        if (c < nchunks) chunk_off[b * nchunks + c] = base + incl - v;
        __syncthreads();
        if (t == 1023) carry = base + incl;
        __syncthreads();
    }
    if (t == 0) n_cand[b] = carry;
}

// ---------------------------------------------------------------------------------------- K3
__global__ void __launch_bounds__(256)
nms_compact_kernel(const uint32_t* __restrict__ st_key, const uint32_t* __restrict__ st_code,
                   const int* __restrict__ chunk_cnt, const int* __restrict__ chunk_off,
                   uint32_t* __restrict__ cp_key, uint32_t* __restrict__ cp_idx, uint32_t* __restrict__ cp_code,
                   int per_chunk_cap, long long cap) {
    const int chunk = blockIdx.x, b = blockIdx.y, nchunks = gridDim.x;
    const int cnt = chunk_cnt[b * nchunks + chunk];
    const int off = chunk_off[b * nchunks + chunk];
    const long long src0 = ((long long)b * nchunks + chunk) * per_chunk_cap;
    const long long dst0 = (long long)b * cap + off;
    for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
        cp_key[dst0 + e] = st_key[src0 + e];
        cp_code[dst0 + e] = st_code[src0 + e];
        cp_idx[dst0 + e] = (uint32_t)(off + e);
    }
}

// ---------------------------------------------------------------------------------------- K4
// Stable LSD radix sort (8-bit digits) of one image's (key, idx) pairs by one CTA.
__global__ void __launch_bounds__(SORT_THREADS)
nms_sort_kernel(uint32_t* __restrict__ key_a, uint32_t* __restrict__ val_a, uint32_t* __restrict__ key_b,
                uint32_t* __restrict__ val_b, const int* __restrict__ n_cand, long long cap) {
    __shared__ int hist[256];
    __shared__ int warp_cnt[SORT_THREADS / 32][256];
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int n = n_cand[b];
    if (n <= 1) return;
    uint32_t* ka = key_a + (long long)b * cap;
    uint32_t* va = val_a + (long long)b * cap;
    uint32_t* kb = key_b + (long long)b * cap;
    uint32_t* vb = val_b + (long long)b * cap;
    for (int pass = 0; pass < 4; ++pass) {
        const int sh = pass * 8;
        if (t < 256) hist[t] = 0;
        __syncthreads();
        for (int e = t; e < n; e += SORT_THREADS) atomicAdd(&hist[(ka[e] >> sh) & 255], 1);
        __syncthreads();
        if (wid == 0) {  // exclusive scan of 256 bins by one warp (8 bins per lane)
            int loc[8], s = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { loc[q] = hist[lane * 8 + q]; s += loc[q]; }
            int incl = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += u;
            }
            int run = incl - s;
#pragma unroll
            for (int q = 0; q < 8; ++q) { hist[lane * 8 + q] = run; run += loc[q]; }
        }
        __syncthreads();
        for (int e0 = 0; e0 < n; e0 += SORT_THREADS) {
            for (int q = t; q < (SORT_THREADS / 32) * 256; q += SORT_THREADS) (&warp_cnt[0][0])[q] = 0;
            __syncthreads();
            const int e = e0 + t;
            const bool has = e < n;
            uint32_t k = 0, v = 0;
            int digit = 0, rank = 0;
            if (has) { k = ka[e]; v = va[e]; digit = (k >> sh) & 255; }
            const unsigned act = __ballot_sync(0xffffffffu, has);
            if (has) {
                const unsigned peers = __match_any_sync(act, digit);
                rank = __popc(peers & ((1u << lane) - 1u));
                if (rank == 0) warp_cnt[wid][digit] = __popc(peers);
            }
            __syncthreads();
            if (t < 256) {  // exclusive scan across warps for digit t; advance the running base
                int run = hist[t];
                for (int wq = 0; wq < SORT_THREADS / 32; ++wq) {
                    const int c = warp_cnt[wq][t];
                    warp_cnt[wq][t] = run;
                    run += c;
                }
                hist[t] = run;
            }
            __syncthreads();
            if (has) {
                const int dst = warp_cnt[wid][digit] + rank;
                kb[dst] = k;
                vb[dst] = v;
            }
            __syncthreads();
        }
        uint32_t* tk = ka; ka = kb; kb = tk;
        uint32_t* tv = va; va = vb; vb = tv;
        __threadfence_block();
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------- K5
__global__ void __launch_bounds__(256)
nms_gather_kernel(const float* __restrict__ pred, const uint32_t* __restrict__ cp_idx, const uint32_t* __restrict__ cp_code,
                  const int* __restrict__ n_cand, float4* __restrict__ box, float4* __restrict__ box_off,
                  float* __restrict__ score, int* __restrict__ cls, int R, int nc, long long cap, int max_nms,
                  float class_offset) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = min(n_cand[b], max_nms);
    if (i >= n) return;
    const uint32_t orig = cp_idx[(long long)b * cap + i];
    const uint32_t code = cp_code[(long long)b * cap + orig];
    const int row = code / nc, j = code - row * nc;
    const float* p = pred + ((long long)b * R + row) * (5 + nc);
    const float cx = p[0], cy = p[1], hw = __fdiv_rn(p[2], 2.f), hh = __fdiv_rn(p[3], 2.f);   // general.py:269-276
    const float4 bx = make_float4(__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh));
    const float c = __fmul_rn((float)j, class_offset);                                         // :494
    const long long o = (long long)b * max_nms + i;
    box[o] = bx;
    box_off[o] = make_float4(__fadd_rn(bx.x, c), __fadd_rn(bx.y, c), __fadd_rn(bx.z, c), __fadd_rn(bx.w, c));
    score[o] = __fmul_rn(p[5 + j], p[4]);
    cls[o] = j;
}

// ---------------------------------------------------------------------------------------- K6
__device__ __forceinline__ float box_area(const float4& a) {
    return __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
}
// inter / (area_a + area_b - inter), fp32, torchvision / box_iou operation order
__device__ __forceinline__ float iou_rn(const float4& a, float area_a, const float4& b, float area_b) {
    const float w = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.f);
    const float h = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.f);
    const float inter = __fmul_rn(w, h);
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
}

__global__ void __launch_bounds__(NMS_THREADS)
nms_suppress_kernel(const float4* __restrict__ box, const float4* __restrict__ box_off, const float* __restrict__ score,
                    const int* __restrict__ cls, const uint32_t* __restrict__ cp_idx, const int* __restrict__ n_cand,
                    float* __restrict__ out, int* __restrict__ counts, int* __restrict__ keep_idx,
                    long long cap, int max_nms, int max_det, float thr_nms, float thr_merge, int merge, int redundant) {
    extern __shared__ __align__(16) unsigned char smraw[];
    float4* kept_box = reinterpret_cast<float4*>(smraw);                       // [max_det]
    float4* ch_box = kept_box + max_det;                                        // [NMS_THREADS]
    float4* merged = ch_box + NMS_THREADS;                                      // [max_det]
    float* kept_area = reinterpret_cast<float*>(merged + max_det);              // [max_det]
    float* ch_area = kept_area + max_det;                                       // [NMS_THREADS]
    int* kept_pos = reinterpret_cast<int*>(ch_area + NMS_THREADS);              // [max_det] sorted position
    int* hits = kept_pos + max_det;                                             // [max_det]
    uint32_t* mask = reinterpret_cast<uint32_t*>(hits + max_det);               // [NMS_THREADS][NMS_WORDS]
    uint32_t* alive_w = mask + NMS_THREADS * NMS_WORDS;                         // [NMS_WORDS]
    __shared__ int s_kept;

    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int n_all = n_cand[b];
    const int n = min(n_all, max_nms);
    const float4* bo = box_off + (long long)b * max_nms;
    if (t == 0) s_kept = 0;
    __syncthreads();

    for (int c0 = 0; c0 < n; c0 += NMS_THREADS) {
        const int kept_before = s_kept;
        if (kept_before >= max_det) break;
        const int i = c0 + t;
        const bool valid = i < n;
        float4 bi = make_float4(0.f, 0.f, 0.f, 0.f);
        float ai = 0.f;
        if (valid) { bi = bo[i]; ai = box_area(bi); }
        ch_box[t] = bi;
        ch_area[t] = ai;
        bool alive = valid;
        for (int k = 0; k < kept_before && alive; ++k)
            alive = !(iou_rn(kept_box[k], kept_area[k], bi, ai) > thr_nms);
        const unsigned aw = __ballot_sync(0xffffffffu, alive);
        if (lane == 0) alive_w[wid] = aw;
        __syncthreads();
        // row t of the intra-chunk suppression matrix: bit j set iff j > t and IoU(t, j) > thr
        for (int wq = 0; wq < NMS_WORDS; ++wq) {
            uint32_t bits = 0;
            if (alive && wq >= wid) {
                const uint32_t cand = alive_w[wq];
                for (int q = 0; q < 32; ++q) {
                    const int j = wq * 32 + q;
                    if (j > t && ((cand >> q) & 1u) && iou_rn(bi, ai, ch_box[j], ch_area[j]) > thr_nms) bits |= 1u << q;
                }
            }
            mask[t * NMS_WORDS + wq] = bits;
        }
        __syncthreads();
        if (wid == 0) {  // sequential greedy resolution by one warp; lane w owns word w of the removed set
            uint32_t removed = 0;
            const uint32_t mine = lane < NMS_WORDS ? alive_w[lane] : 0u;
            int kept = kept_before;
            for (int j = 0; j < NMS_THREADS && kept < max_det; ++j) {
                const int wj = j >> 5;
                const uint32_t aw_j = __shfl_sync(0xffffffffu, mine, wj);
                const uint32_t rm_j = __shfl_sync(0xffffffffu, removed, wj);
                if (((aw_j & ~rm_j) >> (j & 31)) & 1u) {
                    if (lane == 0) { kept_box[kept] = ch_box[j]; kept_area[kept] = ch_area[j]; kept_pos[kept] = c0 + j; }
                    ++kept;
                    if (lane < NMS_WORDS) removed |= mask[j * NMS_WORDS + lane];
                }
            }
            if (lane == 0) s_kept = kept;
        }
        __syncthreads();
    }

    const int K = s_kept;
    const bool do_merge = merge && n_all > 1 && n_all < 3000;            // general.py:499
    const float4* bx = box + (long long)b * max_nms;
    const float* sc = score + (long long)b * max_nms;
    if (do_merge) {
        for (int k = wid; k < K; k += NMS_THREADS / 32) {
            const float4 kb = kept_box[k];
            const float ka = kept_area[k];
            float ax = 0.f, ay = 0.f, az = 0.f, aw2 = 0.f, sw = 0.f;
            int cnt = 0;
            for (int j = lane; j < n; j += 32) {
                const float4 oj = bo[j];
                if (iou_rn(kb, ka, oj, box_area(oj)) > thr_merge) {   // box_iou(boxes[i], boxes) > iou_thres, :501
                    const float w = sc[j];
                    const float4 xj = bx[j];
                    ax = fmaf(w, xj.x, ax); ay = fmaf(w, xj.y, ay); az = fmaf(w, xj.z, az); aw2 = fmaf(w, xj.w, aw2);
                    sw += w;
                    ++cnt;
                }
            }
            ax = warp_sum(ax); ay = warp_sum(ay); az = warp_sum(az); aw2 = warp_sum(aw2); sw = warp_sum(sw);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            if (lane == 0) {
                merged[k] = make_float4(ax / sw, ay / sw, az / sw, aw2 / sw);   // :503
                hits[k] = cnt;
            }
        }
    }
    __syncthreads();
    // ordered output by warp 0
    float* ob = out + (long long)b * max_det * 6;
    int* kb_out = keep_idx ? keep_idx + (long long)b * max_det : nullptr;
    int written = 0;
    if (wid == 0) {
        for (int k0 = 0; k0 < K; k0 += 32) {
            const int k = k0 + lane;
            const bool in = k < K;
            const bool pass = in && (!do_merge || !redundant || hits[k] > 1);   // :504-505
            const unsigned pm = __ballot_sync(0xffffffffu, pass);
            if (pass) {
                const int dst = written + __popc(pm & ((1u << lane) - 1u));
                const int pos = kept_pos[k];
                const float4 v = do_merge ? merged[k] : bx[pos];
                float* o = ob + dst * 6;
                o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
                o[4] = sc[pos];
                o[5] = (float)cls[(long long)b * max_nms + pos];
                if (kb_out) kb_out[dst] = n_all > max_nms ? pos : (int)cp_idx[(long long)b * cap + pos];
            }
            written += __popc(pm);
        }
        if (lane == 0) { counts[b] = written; s_kept = written; }
    }
    __syncthreads();
    const int wr = s_kept;
    for (int e = wr * 6 + t; e < max_det * 6; e += NMS_THREADS) ob[e] = 0.f;
    if (kb_out)
        for (int e = wr + t; e < max_det; e += NMS_THREADS) kb_out[e] = -1;
}

size_t suppress_smem(int max_det) {
    return (size_t)max_det * (2 * sizeof(float4) + sizeof(float) + 2 * sizeof(int)) +
           (size_t)NMS_THREADS * (sizeof(float4) + sizeof(float)) + (size_t)NMS_THREADS * NMS_WORDS * sizeof(uint32_t) +
           NMS_WORDS * sizeof(uint32_t);
}

}  // namespace
}  // namespace sodt

extern "C" size_t sodt_nms_workspace_bytes(int B, int R, int nc, int multi_label) {
    if (B <= 0 || R <= 0 || nc <= 0) return 0;
    return sodt::carve(nullptr, B, R, nc, multi_label && nc > 1, 30000).bytes;
}

extern "C" int sodt_nms(const float* pred, const int* classes, int n_classes, float* out, int* counts, int* keep_idx,
                        void* workspace, size_t workspace_bytes, int B, int R, int nc,
                        float conf_thres, double iou_thres, int multi_label, int agnostic, int merge, int redundant,
                        int max_det, int max_nms, float max_wh, void* stream) {
    using namespace sodt;
    if (!pred || !out || !counts) return SODT_ERR_INVALID_ARG;
    if (B <= 0 || R <= 0 || nc <= 0 || max_det <= 0 || max_nms <= 0) return SODT_ERR_INVALID_ARG;
    if (classes == nullptr) n_classes = 0;
    if ((long long)R * nc > 4294967295LL || max_nms > 30000 || B > 65535) return SODT_ERR_UNSUPPORTED;
    multi_label = multi_label && nc > 1;                           // general.py:441
    if (!workspace || !aligned16(workspace)) return SODT_ERR_WORKSPACE;
    Workspace w = carve(workspace, B, R, nc, multi_label, max_nms);
    if (w.bytes > workspace_bytes) return SODT_ERR_WORKSPACE;
    const size_t sm6 = suppress_smem(max_det);
    if (sm6 > 200 * 1024) return SODT_ERR_UNSUPPORTED;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int no = 5 + nc;
    const int nchunks = (R + CHUNK - 1) / CHUNK;
    const int per = multi_label ? nc : 1;
    const long long cap = (long long)nchunks * CHUNK * per;
    const size_t sm1 = (size_t)CHUNK * no * sizeof(float);
    if (sm1 > 200 * 1024) return SODT_ERR_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1);
    if (e != cudaSuccess) return cuda_status(e);
    e = cudaFuncSetAttribute(nms_suppress_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm6);
    if (e != cudaSuccess) return cuda_status(e);
    const int vec_ok = aligned16(pred) && ((long long)R * no) % 4 == 0 && (CHUNK * no) % 4 == 0;
    int st;
    nms_scan_kernel<<<dim3(nchunks, B), CHUNK, sm1, s>>>(pred, classes, n_classes, w.st_key, w.st_code, w.chunk_cnt, R, nc,
                                                         conf_thres, multi_label, vec_ok);
    if ((st = check_launch()) != SODT_OK) return st;
    nms_offsets_kernel<<<B, 1024, 0, s>>>(w.chunk_cnt, w.chunk_off, w.n_cand, nchunks);
    if ((st = check_launch()) != SODT_OK) return st;
    nms_compact_kernel<<<dim3(nchunks, B), 256, 0, s>>>(w.st_key, w.st_code, w.chunk_cnt, w.chunk_off, w.cp_key, w.cp_idx,
                                                        w.cp_code, CHUNK * per, cap);
    if ((st = check_launch()) != SODT_OK) return st;
    nms_sort_kernel<<<B, SORT_THREADS, 0, s>>>(w.cp_key, w.cp_idx, w.st_key, w.st_code, w.n_cand, cap);
    if ((st = check_launch()) != SODT_OK) return st;
    nms_gather_kernel<<<dim3((max_nms + 255) / 256, B), 256, 0, s>>>(pred, w.cp_idx, w.cp_code, w.n_cand, w.box, w.box_off,
                                                                     w.score, w.cls, R, nc, cap, max_nms,
                                                                     agnostic ? 0.f : max_wh);
    if ((st = check_launch()) != SODT_OK) return st;
    // torchvision compares the fp32 IoU with the threshold as a double; box_iou(...) > iou_thres compares in fp32.
    const float t32 = (float)iou_thres;
    const float thr_nms = ((double)t32 > iou_thres) ? nextafterf(t32, -INFINITY) : t32;
    nms_suppress_kernel<<<B, NMS_THREADS, sm6, s>>>(w.box, w.box_off, w.score, w.cls, w.cp_idx, w.n_cand, out, counts, keep_idx,
                                                    cap, max_nms, max_det, thr_nms, t32, merge, redundant);
    return check_launch();
}
