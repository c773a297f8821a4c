// hostcheck.cpp -- TEST-ONLY.  Compiles the product's host/device core header (tb_core.cuh) with g++
// so the bit-parallel math the CUDA kernels execute can be checked against the oracle on a machine
// without a GPU.  Nothing in tetris_b200 loads this library.
#include <cstdint>
#include <cstring>
#include "../../tetris_b200/csrc/tb_core.cuh"

using namespace tb;

template <int C, int R>
static int afterstates_t(int piece, const uint16_t *rows, int mode, float *feats, uint8_t *terminal,
                         int32_t *ncleared, uint16_t *rows_out, int32_t *anchor, uint8_t *used_fast)
{
    using S = Shape<C, R>;
    uint32_t w[S::NW];
    std::memset(w, 0, sizeof w);
    for (int r = 0; r < S::N; ++r) w[r >> 1] |= (uint32_t)rows[r] << (16 * (r & 1));
    uint32_t col[C];
    rows_to_cols<C, R>(w, col);
    static uint32_t runtab[RunTab<R>::WORDS];
    static bool init = false;
    if (!init) {
        for (int m = 0; m < RunTab<R>::SIZE; ++m) {
            runtab[m] = run_tab_entry<R>((uint32_t)m);
            if (RunTab<R>::COPIES == 2) runtab[RunTab<R>::SIZE + m] = prmt(runtab[m], 0u, 0x0423);   // the permuted copy
        }
        init = true;
    }
    uint32_t rec[Env<C, R>::WORDS + 8];
    std::memset(rec, 0, sizeof rec);
    uint32_t any = 0;
    for (int k = 0; k < C; ++k) any |= col[k];
    const int hmax = height_of(any);
    if (hmax <= R) {                                        // the record's precondition: a non-terminal board
        for (int k = 0; k < C; ++k) rec[Env<C, R>::COLX + 2 + k] = col[k];
        build_env<C, R>(runtab, rec);
    }
    const uint32_t pw = kPieceHost[piece];
    const int n = piece_num_slots(pw, C);
    const unsigned long long vslots = hmax <= R ? valid_slots<C, R>(col, pw, kOriHost) : 0ull;
    if (hmax <= R && (valid_slots<C, R, true>(col, pw, kOriHost) != 0ull) != (vslots != 0ull)) return -9;
    for (int s = 0; s < n; ++s) {
        int ori, c;
        slot_to_placement(pw, C, s, ori, c);
        const uint32_t d = kOriHost[ori];
        Eval quick, full;
        uint32_t nc[C];
        // mode 0: the incremental path the kernels use, falling back to the general path exactly where they do
        const int status = (mode == 0 && hmax <= R) ? eval_placement_w<C, R>(rec, runtab, d, c, quick) : -1;
        eval_slow<C, R>(col, d, c, full, nc);                      // also yields the afterstate board
        const bool fast = status == kFastDone;
        if (status >= 0) {
            // what eval_placement reports on its early exits must agree with the general path
            if (quick.a != full.a || quick.terminal != full.terminal) return -3;
            if (status != kFastTerminal && quick.full != full.full) return -4;
            if (status == kFastTerminal && !(full.terminal && full.full == 0u)) return -5;
            if (status == kFastClears && full.full == 0u) return -6;
        }
        const Eval &e = fast ? quick : full;
        std::memcpy(feats + 8 * s, e.f, sizeof e.f);
        terminal[s] = (uint8_t)full.terminal;
        ncleared[s] = popc32(full.full);
        anchor[s] = full.a;
        used_fast[s] = fast;
        uint32_t wo[S::NW];
        cols_to_rows<C, R>(nc, wo);
        for (int r = 0; r < S::N; ++r) rows_out[s * S::N + r] = (uint16_t)(wo[r >> 1] >> (16 * (r & 1)));
        // placement_valid (heights + full-row count only) must agree with the terminal flag
        if (hmax <= R && placement_valid<C, R>(col, d, c, hmax) != (full.terminal == 0)) return -1;
        if (hmax <= R && (int)((vslots >> s) & 1ull) != (full.terminal == 0)) return -7;   // heights + near-full gate
    }
    return n;
}

#define SHAPES(X) X(10, 20) X(10, 10) X(6, 12) X(8, 16) X(4, 4) X(16, 27) X(12, 24) X(7, 9)

extern "C" int hc_afterstates(int C, int R, int piece, const uint16_t *rows, int mode, float *feats,
                              uint8_t *terminal, int32_t *ncleared, uint16_t *rows_out, int32_t *anchor,
                              uint8_t *used_fast)
{
#define X(c, r) if (C == c && R == r) return afterstates_t<c, r>(piece, rows, mode, feats, terminal, ncleared, rows_out, anchor, used_fast);
    SHAPES(X)
#undef X
    return -2;
}

template <int C, int R>
static int transpose_t(const uint16_t *rows, uint16_t *rows_back, uint32_t *cols_out)
{
    using S = Shape<C, R>;
    uint32_t w[S::NW], w2[S::NW], col[C];
    std::memset(w, 0, sizeof w);
    for (int r = 0; r < S::N; ++r) w[r >> 1] |= (uint32_t)rows[r] << (16 * (r & 1));
    rows_to_cols<C, R>(w, col);
    for (int c = 0; c < C; ++c) cols_out[c] = col[c];
    cols_to_rows<C, R>(col, w2);
    for (int r = 0; r < S::N; ++r) rows_back[r] = (uint16_t)(w2[r >> 1] >> (16 * (r & 1)));
    return 0;
}
extern "C" int hc_transpose(int C, int R, const uint16_t *rows, uint16_t *rows_back, uint32_t *cols_out)
{
#define X(c, r) if (C == c && R == r) return transpose_t<c, r>(rows, rows_back, cols_out);
    SHAPES(X)
#undef X
    return -2;
}

extern "C" uint32_t hc_rng(uint64_t seed, uint64_t env, uint32_t ctr, uint32_t stream)
{
    return rng32(env_key(seed, env), ctr, stream);
}
extern "C" int hc_bag_draw(int n_set, uint64_t seed, uint64_t env, uint32_t *bag, uint32_t *draws)
{
    return bag_draw(n_set, env_key(seed, env), *bag, *draws);
}
extern "C" float hc_fitness(const float *f, const float *w) { return fitness(f, w); }
extern "C" int hc_num_slots(int piece, int C) { return piece_num_slots(kPieceHost[piece], C); }

// The compile-time table images the kernels copy must equal what decode_ori / run_tab_entry compute.
template <int R>
static int run_image_ok()
{
    static constexpr RunImage<R> img = make_run_image<R>();
    for (int m = 0; m < RunTab<R>::SIZE; ++m) {
        if (img.v[m] != run_tab_entry<R>((uint32_t)m)) return 0;
        if (RunTab<R>::COPIES == 2 && img.v[RunTab<R>::SIZE + m] != prmt(run_tab_entry<R>((uint32_t)m), 0u, 0x0423)) return 0;
    }
    return 1;
}
extern "C" int hc_table_images()
{
    static constexpr OdescImage od = make_odesc_image();
    for (int i = 0; i < kNumOris; ++i) {
        const OriU u = decode_ori(kOriHost[i]);
        uint32_t w[31];
        std::memcpy(w, &u, sizeof w);
        for (int k = 0; k < 31; ++k)
            if (od.w[i][k] != w[k]) return -(100 + i);
        if (od.w[i][31] != 0u) return -(200 + i);
    }
    if (!run_image_ok<20>()) return -20;
    if (!run_image_ok<10>()) return -10;
    if (!run_image_ok<12>()) return -12;
    if (!run_image_ok<16>()) return -16;
    if (!run_image_ok<4>()) return -4;
    if (!run_image_ok<27>()) return -27;
    if (!run_image_ok<24>()) return -24;
    if (!run_image_ok<9>()) return -9;
    return 0;
}

// The bit helpers of the column-parallel legality test and of the branch-free placement, against plain restatements.
extern "C" int hc_bit_helpers()
{
    // interleave16: bit i of the low half -> bit 2i, of the high half -> bit 2i + 1
    uint32_t x = 0x9E3779B9u;
    for (int it = 0; it < 4096; ++it) {
        x = x * 1664525u + 1013904223u;
        uint32_t want = 0u;
        for (int i = 0; i < 16; ++i) want |= ((x >> i) & 1u) << (2 * i) | ((x >> (16 + i)) & 1u) << (2 * i + 1);
        if (interleave16(x) != want) return -1;
        if ((spread16(x & 0xFFFFu) | spread16(x >> 16) << 1) != want) return -2;
    }
    for (int i = 0; i < kNumOris; ++i) {
        const uint32_t d = kOriHost[i];
        // poke_sel: the selector of piece column dx picks the 16-bit field k = 4 - ph + bot[dx] of a 64-bit word
        const uint32_t lo = 0x22221111u, hi = 0x44443333u;            // field k holds 0x1111 * (k + 1)
        for (int dx = 0; dx < desc_w(d); ++dx) {
            const int k = 4 - desc_ph(d) + desc_bot(d, dx);
            if (k < 0 || k > 3) return -(100 + i);
            if ((prmt(lo, hi, (uint32_t)(poke_sel(i) >> (16 * dx))) & 0xFFFFu) != 0x1111u * (uint32_t)(k + 1)) return -(200 + i);
        }
        // piece_cells4: field dx = the column's cells relative to the anchor row
        for (int dx = 0; dx < 4; ++dx) {
            const uint32_t want = desc_len(d, dx) ? mask_lo(desc_len(d, dx)) << desc_bot(d, dx) : 0u;
            if (((piece_cells4(d) >> (4 * dx)) & 15u) != want) return -(300 + i);
        }
    }
    // hole_depth_of against the plain loop
    for (int it = 0; it < 4096; ++it) {
        x = x * 1664525u + 1013904223u;
        const uint32_t col = x & 0x00FFFFFFu;
        const int h = height_of(col);
        const uint32_t hole = ~col & mask_lo(h), t = hole & (col >> 1);
        int want = 0;
        for (uint32_t tt = t; tt; tt &= tt - 1u) want += popc32(col >> (ctz32(tt) + 1));
        if (hole_depth_of(col, t) != want) return -3;
    }
    return 0;
}
