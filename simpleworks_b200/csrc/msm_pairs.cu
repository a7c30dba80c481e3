// MSM stage 3b: batch-affine pair sums (fewer field products per point than the XYZZ accumulation).
//
// After the sort, neighbouring positions mostly hold points of the same bucket (a bucket has tens to hundreds
// of them).  Level 1 adds the points of positions (2i, 2i+1) when both carry the same valid bucket id, level 2
// adds the level-1 sums of slots (2i, 2i+1) -- i.e. four positions of one bucket -- and so on, in AFFINE
// coordinates:
//     lambda = (y2 - y1) / (x2 - x1),   x3 = lambda^2 - x1 - x2,   y3 = lambda (x1 - x3) - y1.
// k_msm_accumulate_paired then adds ONE point per summed block of 2^L positions, so after L levels only about
// n / 2^L of the 10-product XYZZ mixed additions are left; a pair sum costs 6 products (1 forward, 5 backward)
// once the division is shared: Montgomery's trick turns the inversions of all denominators of a tile of
// 128 x M slots into ONE inversion plus 3 products per denominator.
//
// Per level and chunk of slots, three kernels:
//   k_pair_fwd   gathers the two points of every slot (level 1: table points selected by the sorted values, with
//                their signs; higher levels: the two sums of the level below, which lie next to each other),
//                decides whether the slot can be summed, and streams out dx, dy, x1, y1 and the running product of
//                the thread's denominators before this slot; the 128 running products of a block are multiplied
//                in a shared-memory tree -> one product per tile.           (light: 1 product per slot; all gathers)
//   k_pair_inv   one thread per tile inverts its product with the binary extended Euclidean algorithm (ALU pipe).
//   k_pair_bwd   rebuilds the tree, walks it down with the inverse, and every thread unwinds its prefixes from the
//                last slot to the first, computes the sums and writes R_L[slot] (or a marker) and the level map.
//                                                                  (5 products per slot, purely streaming accesses)
// A slot that cannot be summed -- different buckets, a zero digit, an identity base, equal or opposite points
// (x2 == x1), a child that was not summed -- is marked; the consumer then falls back to the level below, down to
// the table points themselves, so nothing is assumed about the bases.
//
// Thread t of tile b handles slots chunk0 + (b M + j) 128 + t (j < M): consecutive threads touch consecutive
// slots, so keys, values, the streamed records (one plane per field element) and R_L are all accessed coalesced;
// only level 1's table points are gathered.
#include "msm_common.cuh"

namespace swb {

constexpr int PAIR_THREADS = 128;
#ifndef PAIR_FWD
#define PAIR_FWD 2                 // slots whose loads are in flight together in the forward pass (M is a multiple)
#endif
#ifndef PAIR_FWD_BLOCKS
#define PAIR_FWD_BLOCKS 3
#endif
constexpr uint32_t PAIR_MARK = 0xffffffffu;      // top limb of x of a slot that was not summed (no field element has it)

__device__ __forceinline__ Fq ld_fq(const Fq* __restrict__ p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    Fq r;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const uint4 a = q[k];
        r.l[4 * k] = a.x; r.l[4 * k + 1] = a.y; r.l[4 * k + 2] = a.z; r.l[4 * k + 3] = a.w;
    }
    return r;
}
__device__ __forceinline__ void st_fq(Fq* __restrict__ p, const Fq& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int k = 0; k < 3; k++) q[k] = make_uint4(v.l[4 * k], v.l[4 * k + 1], v.l[4 * k + 2], v.l[4 * k + 3]);
}

struct PairScratch {        // per chunk, one plane per streamed value, indexed by slot - chunk0
    Fq *pre, *dx, *dy, *x1, *y1;
    uint8_t* flag;          // slot is being summed
    Fq* leaf;               // [tiles][128] the threads' running products
    Fq* tile_prod;          // [tiles] product of a tile's denominators, inverted in place by k_pair_inv
};

// product tree over the 128 running products of a block (leaves in the upper half); returns with tree[1] = root
__device__ __forceinline__ void tree_up(Fq* tree, uint32_t t) {
    __syncthreads();
    for (uint32_t w = PAIR_THREADS / 2; w >= 1; w >>= 1) {
        if (t < w) tree[w + t] = tree[2 * (w + t)] * tree[2 * (w + t) + 1];
        __syncthreads();
    }
}

// the two points of slot s at level L (first positions p and pm): false when the slot cannot be summed
template <bool FROM_TABLE>
__device__ __forceinline__ bool pair_load(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, const Fq* __restrict__ src,
                                          size_t total, size_t n, uint32_t B, size_t s, size_t p, size_t pm, Fq& x1, Fq& y1, Fq& dx,
                                          Fq& dy) {
    if (pm >= total) return false;
    const uint32_t k0 = keys[p], k1 = keys[pm];
    // same valid bucket and same bucket set (positions fit 32 bits; on the table path n == total)
    if (!(k0 == k1 && k0 < B && (n >= total || ((uint32_t)pm % (uint32_t)n) != 0u))) return false;
    Fq x2, y2;
    if (FROM_TABLE) {
        const uint32_t w0 = vals[p], w1 = vals[pm];
        const Fq* a = src + 2 * (size_t)(w0 & 0x7fffffffu);
        const Fq* b = src + 2 * (size_t)(w1 & 0x7fffffffu);
        x1 = ld_fq(a); y1 = ld_fq(a + 1);
        x2 = ld_fq(b); y2 = ld_fq(b + 1);
        // an identity base is stored as (0, 0) (x = 0 alone is not enough: (0, +-1) lies on the curve)
        if ((x1.is_zero() && y1.is_zero()) || (x2.is_zero() && y2.is_zero())) return false;
        if (w0 >> 31) y1 = y1.neg();
        if (w1 >> 31) y2 = y2.neg();
    } else {
        const Fq* a = src + 4 * s;                        // children 2s and 2s + 1 of the level below
        x1 = ld_fq(a); x2 = ld_fq(a + 2);
        if (x1.l[11] == PAIR_MARK || x2.l[11] == PAIR_MARK) return false;
        y1 = ld_fq(a + 1); y2 = ld_fq(a + 3);
    }
    dx = x2 - x1;
    dy = y2 - y1;
    return !dx.is_zero();                                 // equal or opposite points: left to the XYZZ addition
}

// Level 1 (FROM_TABLE) streams dx, dy, x1, y1 out for k_pair_bwd, which must not repeat the gathers; higher levels only
// the prefix products -- their inputs lie contiguously in the level below and are simply read again.
template <bool FROM_TABLE>
__global__ void __launch_bounds__(PAIR_THREADS, PAIR_FWD_BLOCKS) k_pair_fwd(PairScratch sc, const uint32_t* __restrict__ keys,
                                                              const uint32_t* __restrict__ vals, const Fq* __restrict__ src,
                                                              size_t total, size_t n, uint32_t B, uint32_t L, size_t chunk0,
                                                              size_t chunk_slots, uint32_t M) {
    __shared__ Fq tree[2 * PAIR_THREADS];
    const uint32_t t = threadIdx.x;
    const size_t tile0 = (size_t)blockIdx.x * M * PAIR_THREADS;      // first slot of the tile inside the chunk
    const size_t half = (size_t)1 << (L - 1);
    Fq acc = Fq::one();
    // PAIR_FWD slots per iteration: all their loads are issued before the first product, so the gathers of several
    // slots are in flight at once (the pass is bound by the latency of the gathers, not by arithmetic)
    for (uint32_t j0 = 0; j0 < M; j0 += PAIR_FWD) {
        Fq x1[PAIR_FWD], y1[PAIR_FWD], dx[PAIR_FWD], dy[PAIR_FWD];
        bool ok[PAIR_FWD];
        size_t ls[PAIR_FWD];
#pragma unroll
        for (int q = 0; q < PAIR_FWD; q++) {
            ls[q] = tile0 + (size_t)(j0 + q) * PAIR_THREADS + t;      // slot inside the chunk
            ok[q] = false;
            if (ls[q] < chunk_slots) {
                const size_t s = chunk0 + ls[q];
                const size_t p = s << L;
                ok[q] = pair_load<FROM_TABLE>(keys, vals, src, total, n, B, s, p, p + half, x1[q], y1[q], dx[q], dy[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < PAIR_FWD; q++) {
            if (ls[q] >= chunk_slots) continue;
            sc.flag[ls[q]] = ok[q] ? 1 : 0;
            if (ok[q]) {
                st_fq(sc.pre + ls[q], acc);
                if (FROM_TABLE) {
                    st_fq(sc.dx + ls[q], dx[q]);
                    st_fq(sc.dy + ls[q], dy[q]);
                    st_fq(sc.x1 + ls[q], x1[q]);
                    st_fq(sc.y1 + ls[q], y1[q]);
                }
                acc = acc * dx[q];
            }
        }
    }
    st_fq(sc.leaf + (size_t)blockIdx.x * PAIR_THREADS + t, acc);
    tree[PAIR_THREADS + t] = acc;
    tree_up(tree, t);
    if (t == 0) st_fq(sc.tile_prod + blockIdx.x, tree[1]);
}

__global__ void __launch_bounds__(64) k_pair_inv(Fq* __restrict__ tile_prod, uint32_t tiles) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= tiles) return;
    st_fq(tile_prod + i, ld_fq(tile_prod + i).inverse_bingcd());
}

template <bool RECORDS>
__global__ void __launch_bounds__(PAIR_THREADS, 4) k_pair_bwd(Fq* __restrict__ R, uint8_t* __restrict__ lvl, PairScratch sc,
                                                              const Fq* __restrict__ below, uint32_t L, size_t chunk0,
                                                              size_t chunk_slots, uint32_t M, uint32_t* __restrict__ summed) {
    __shared__ Fq tree[2 * PAIR_THREADS];
    const uint32_t t = threadIdx.x;
    const size_t tile0 = (size_t)blockIdx.x * M * PAIR_THREADS;
    tree[PAIR_THREADS + t] = ld_fq(sc.leaf + (size_t)blockIdx.x * PAIR_THREADS + t);
    tree_up(tree, t);
    if (t == 0) tree[1] = ld_fq(sc.tile_prod + blockIdx.x);            // the inverted root
    __syncthreads();
    for (uint32_t w = 1; w < PAIR_THREADS; w <<= 1) {
        if (t < w) {
            const Fq inv = tree[w + t], l = tree[2 * (w + t)], r = tree[2 * (w + t) + 1];
            tree[2 * (w + t)] = inv * r;
            tree[2 * (w + t) + 1] = inv * l;
        }
        __syncthreads();
    }
    Fq inv = tree[PAIR_THREADS + t];                                    // 1 / (product of this thread's denominators)
    uint32_t mine = 0;
    for (uint32_t j = M; j-- > 0;) {
        const size_t ls = tile0 + (size_t)j * PAIR_THREADS + t;
        if (ls >= chunk_slots) continue;
        const size_t s = chunk0 + ls;
        Fq* out = R + 2 * s;
        if (!sc.flag[ls]) {
            // only the top limbs of x need to say "not summed"
            reinterpret_cast<uint4*>(out)[2] = make_uint4(PAIR_MARK, PAIR_MARK, PAIR_MARK, PAIR_MARK);
            continue;
        }
        Fq x1, dx;
        if (RECORDS) {
            dx = ld_fq(sc.dx + ls);
            x1 = ld_fq(sc.x1 + ls);
        } else {
            x1 = ld_fq(below + 4 * s);
            dx = ld_fq(below + 4 * s + 2) - x1;
        }
        Fq lam = inv * ld_fq(sc.pre + ls);                              // 1 / dx
        inv = inv * dx;
        const Fq y1 = RECORDS ? ld_fq(sc.y1 + ls) : ld_fq(below + 4 * s + 1);
        lam = (RECORDS ? ld_fq(sc.dy + ls) : ld_fq(below + 4 * s + 3) - y1) * lam;
        const Fq x3 = lam.sqr() - x1 - x1 - dx;                         // x2 = x1 + dx
        st_fq(out, x3);
        st_fq(out + 1, lam * (x1 - x3) - y1);
        lvl[(s << L) >> 1] = (uint8_t)L;                                 // indexed by the block's first position pair
        mine++;
    }
    if (summed) {                                                       // profiling: how many slots were really summed
        mine = __reduce_add_sync(0xffffffffu, mine);
        if ((t & 31u) == 0 && mine) atomicAdd(summed, mine);
    }
}

// Runs `levels` levels of pair sums over the sorted pairs.  R[l] (l = 1 .. levels) receives the level-l sums
// (ceil(total / 2^l) slots of two field elements), lvl (ceil(total / 2) bytes, zeroed here) the level map:
// lvl[p / 2] = highest level whose summed block STARTS at the even position p (0 = not even the pair (p, p + 1)).
int msm_launch_pair_sums(swb_ctx* c, const MsmPlan& pl, Fq* const* R, uint8_t* lvl, int levels, const uint32_t* sorted_keys,
                         const uint32_t* sorted_vals, const Fq* bases, StageTimer* tm) {
    const size_t total = pl.total;
    if (total < 2 || levels < 1) return SWB_OK;
    SWB_CUDA(c, cudaMemsetAsync(lvl, 0, (total + 1) / 2, c->stream));
    uint32_t* summed = nullptr;
    if (tm && tm->on()) {
        summed = (uint32_t*)get_scratch(c, "msm_pair_count", 64);
        if (!summed) return SWB_ENOMEM;
        SWB_CUDA(c, cudaMemsetAsync(summed, 0, sizeof(uint32_t), c->stream));
    }
    // chunk: bounds the streamed records (5 field elements = 240 bytes per slot)
    const size_t chunk_cap = (size_t)1 << 25;
    for (int L = 1; L <= levels; L++) {
        const size_t nslots = (total + ((size_t)1 << L) - 1) >> L;
        if (nslots == 0) break;
        for (size_t chunk0 = 0; chunk0 < nslots; chunk0 += chunk_cap) {
            const size_t cs = nslots - chunk0 < chunk_cap ? nslots - chunk0 : chunk_cap;
            // M slots per thread: large tiles amortise the tree and the inversion, small ones keep enough tiles in flight
            uint32_t M = 64;
            while (M > 4 && cs / ((size_t)M * PAIR_THREADS) < (size_t)c->sm_count * 8) M >>= 1;
            const size_t per_tile = (size_t)M * PAIR_THREADS;
            const uint32_t tiles = (uint32_t)((cs + per_tile - 1) / per_tile);
            const size_t planes = L == 1 ? 5 : 1;        // level 1 streams five values per slot, the others one
            uint8_t* raw = (uint8_t*)get_scratch(c, "msm_pair_records", cs * (planes * sizeof(Fq) + 1) + (size_t)tiles * (PAIR_THREADS + 1) * sizeof(Fq) + 4096);
            if (!raw) return SWB_ENOMEM;
            PairScratch sc;
            sc.pre = (Fq*)raw;
            sc.dx = sc.pre + cs;
            sc.dy = sc.dx + (L == 1 ? cs : 0);
            sc.x1 = sc.dy + (L == 1 ? cs : 0);
            sc.y1 = sc.x1 + (L == 1 ? cs : 0);
            sc.leaf = sc.pre + planes * cs;
            sc.tile_prod = sc.leaf + (size_t)tiles * PAIR_THREADS;
            sc.flag = (uint8_t*)(sc.tile_prod + tiles);
            if (tm) tm->span_begin(L == 1 ? "pair_fwd_gather" : "pair_fwd_upper");
            if (L == 1)
                k_pair_fwd<true><<<tiles, PAIR_THREADS, 0, c->stream>>>(sc, sorted_keys, sorted_vals, bases, total, pl.seg_len, pl.key_space,
                                                                         (uint32_t)L, chunk0, cs, M);
            else
                k_pair_fwd<false><<<tiles, PAIR_THREADS, 0, c->stream>>>(sc, sorted_keys, sorted_vals, R[L - 1], total, pl.seg_len, pl.key_space,
                                                                          (uint32_t)L, chunk0, cs, M);
            SWB_LAUNCH_CHECK(c, "k_pair_fwd");
            if (tm) tm->span_end();
            k_pair_inv<<<(tiles + 63) / 64, 64, 0, c->stream>>>(sc.tile_prod, tiles);
            SWB_LAUNCH_CHECK(c, "k_pair_inv");
            if (tm) tm->span_begin("pair_bwd");
            if (L == 1) k_pair_bwd<true><<<tiles, PAIR_THREADS, 0, c->stream>>>(R[L], lvl, sc, nullptr, (uint32_t)L, chunk0, cs, M, summed);
            else k_pair_bwd<false><<<tiles, PAIR_THREADS, 0, c->stream>>>(R[L], lvl, sc, R[L - 1], (uint32_t)L, chunk0, cs, M, summed);
            SWB_LAUNCH_CHECK(c, "k_pair_bwd");
            if (tm) tm->span_end();
        }
    }
    if (summed) {
        if (!c->pair_count_host && cudaMallocHost(&c->pair_count_host, 64) != cudaSuccess) c->pair_count_host = nullptr;
        uint32_t* host = c->pair_count_host;
        if (host) {
            SWB_CUDA(c, cudaMemcpyAsync(host, summed, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
            tm->counter("pair_slots_summed_millions", host);
        }
    }
    return SWB_OK;
}

}  // namespace swb
