// cb_plan_pack.h - tile packing of the shell stream plan (cb_internal.h: CbStreamShape / CbTileS), written
// once for host and device: no allocation, fixed-size scratch, no library calls, integer arithmetic only - the
// device-built plan (cb_plan_device.cuh) and the host-built one (cb_api.cu) are the same bits by construction
// and a test compares them.
//
// Input: the joint-pair blocks in natural order (grouped by column joint B ascending, row joint A ascending
// inside), each with its contribution list, and a CSR joint -> first block.  The owned joints are cut into
// SEGMENTS of CB_PK_SEG joints that are packed independently (a tile never crosses a segment boundary), so a
// segment is one unit of parallel work: pass 1 counts its tiles, an exclusive scan places them, pass 2 emits.
// Inside a segment the packing is greedy: a joint joins the open tile if the tile stays a contiguous slice of
// Ax, its image / shell slots / pair records fit the kernel shape and the joint's block parts find room in the
// 32 lanes (parts in decreasing size, first lane with room; a block with more contributions than a lane has
// steps is cut in two, the second part a follower that closes its lane).
#ifndef CB_PLAN_PACK_H
#define CB_PLAN_PACK_H
#include "cb_internal.h"

#ifdef __CUDACC__
#define CB_HD __host__ __device__
#else
#define CB_HD
#endif

#define CB_PK_MAXBLK 96
#define CB_PK_MAXEL 48
#define CB_PK_SEG 512

struct PkPart { uint8_t blk, c0, cnt, follow; };
struct PkLane { uint8_t load, closed, np, pad; PkPart parts[8]; };
struct PkIn {
    const CbPair *pairs; const CbContrib *contribs;
    const int32_t *jpair;       // [NJ+1] first block of joint B (blocks of B: jpair[B] .. jpair[B+1])
    const int32_t *nfree, *colh; const int64_t *base; int64_t ax_base; int ax_pad;
    const int32_t *cls;         // geometry class of every shell, or nullptr
    CbStreamShape shp;
};
struct PkOut { CbTileS *tiles; uint32_t *steps, *pairsS; int32_t *elems, *blk; };
struct PkTile {
    int64_t out0; int32_t nout; int nblk, nel, nlanes;
    int32_t blk[CB_PK_MAXBLK];
    int32_t el[CB_PK_MAXEL];
    PkLane lanes[32];
};

// place the parts of blocks [p0, p1) (tile-local block numbers blk0 ...) into the lanes; false if > 32 lanes
CB_HD inline bool pk_place(const PkIn &in, PkLane *lanes, int &nlanes, int p0, int p1, int blk0)
{
    const int S = in.shp.steps;
    PkPart parts[2 * 16];
    int np = 0;
    if (p1 - p0 > 16) return false;
    for (int q = p0; q < p1; ++q) {
        const int cnt = in.pairs[q].ccount, b = blk0 + (q - p0);
        if (cnt <= S) { PkPart p = {(uint8_t)b, 0, (uint8_t)cnt, 0}; parts[np++] = p; }
        else {
            const int h = (cnt + 1) / 2;
            PkPart a = {(uint8_t)b, 0, (uint8_t)h, 0}, f = {(uint8_t)b, (uint8_t)h, (uint8_t)(cnt - h), 1};
            parts[np++] = a; parts[np++] = f;
        }
    }
    for (int i = 1; i < np; ++i) {                    // stable insertion sort, decreasing size
        const PkPart k = parts[i]; int j = i - 1;
        while (j >= 0 && parts[j].cnt < k.cnt) { parts[j + 1] = parts[j]; --j; }
        parts[j + 1] = k;
    }
    for (int i = 0; i < np; ++i) {
        int l = 0;
        for (; l < nlanes; ++l)
            if (!lanes[l].closed && lanes[l].load + parts[i].cnt <= S && lanes[l].np < 8) break;
        if (l == nlanes) {
            if (nlanes == 32) return false;
            lanes[l].load = 0; lanes[l].closed = 0; lanes[l].np = 0; lanes[l].pad = 0; ++nlanes;
        }
        lanes[l].parts[lanes[l].np++] = parts[i]; lanes[l].load += parts[i].cnt;
        if (parts[i].follow) lanes[l].closed = 1;     // a follower is the last part of its lane
    }
    return true;
}

// lane order, shell slots (performance only) and the records of tile ti
CB_HD inline void pk_close(const PkIn &in, PkTile &T, long ti, const PkOut &out)
{
    const int S = in.shp.steps, SL = in.shp.slots, NP = in.shp.pairs;
    int nsteps = 0;
    for (int l = 0; l < T.nlanes; ++l) if (T.lanes[l].load > nsteps) nsteps = T.lanes[l].load;
    // (1) lanes that store at the same steps next to each other; inside such a group the lanes of an aligned
    // octet get blocks whose image offsets fall into distinct 16-byte bank groups, as far as a greedy pick manages
    const int shift = (int)((T.out0 + in.ax_pad) & 1);
    unsigned sig[32]; signed char key[32][CB_S_MAXSTEPS_ANY];
    for (int l = 0; l < T.nlanes; ++l) {
        sig[l] = 0;
        for (int k = 0; k < CB_S_MAXSTEPS_ANY; ++k) key[l][k] = -1;
        int st = 0;
        for (int q = 0; q < T.lanes[l].np; ++q) {
            const PkPart pt = T.lanes[l].parts[q];
            st += pt.cnt;
            const CbPair p = in.pairs[T.blk[pt.blk]];
            const int rel = (int)(p.off - T.out0);
            const bool fast = p.maskA == 0x3f && p.maskB == 0x3f && !((shift + rel) & 1) && !(p.colh & 1);
            if (pt.follow) sig[l] |= 1u << 31;
            else { sig[l] |= 1u << (st - 1); key[l][st - 1] = (signed char)(fast ? ((shift + rel) >> 1) & 7 : -1); }
        }
    }
    int order[32];
    for (int l = 0; l < T.nlanes; ++l) order[l] = l;
    for (int i = 1; i < T.nlanes; ++i) {              // stable insertion sort by signature
        const int k = order[i]; int j = i - 1;
        while (j >= 0 && sig[order[j]] > sig[k]) { order[j + 1] = order[j]; --j; }
        order[j + 1] = k;
    }
    int fin[32]; bool taken[32]; unsigned char used[CB_S_MAXSTEPS_ANY][8];
    for (int l = 0; l < T.nlanes; ++l) taken[l] = false;
    for (int pos = 0; pos < T.nlanes; ++pos) {
        if ((pos & 7) == 0)
            for (int k = 0; k < CB_S_MAXSTEPS_ANY; ++k) for (int b = 0; b < 8; ++b) used[k][b] = 0;
        int head = 0;
        while (taken[head]) ++head;
        const unsigned sg = sig[order[head]];
        int pick = head;
        for (int c = head, seen = 0; c < T.nlanes && sig[order[c]] == sg && seen < 16; ++c) {
            if (taken[c]) continue;
            ++seen;
            bool clash = false;
            for (int k = 0; k < CB_S_MAXSTEPS_ANY; ++k) { const int ky = key[order[c]][k]; if (ky >= 0 && used[k][ky]) clash = true; }
            if (!clash) { pick = c; break; }
        }
        taken[pick] = true; fin[pos] = order[pick];
        for (int k = 0; k < CB_S_MAXSTEPS_ANY; ++k) { const int ky = key[order[pick]][k]; if (ky >= 0) used[k][ky] = 1; }
    }
    // (2) shell slots: the records read by one aligned octet of lanes in one step in distinct 16-byte bank groups,
    // i.e. slot numbers distinct modulo 8 (a record is nine 16-byte units long).  Greedy colouring by conflict count.
    unsigned char el_of[32][CB_S_MAXSTEPS_ANY];       // [final lane][step] -> tile-local shell index (255: idle)
    for (int pos = 0; pos < T.nlanes; ++pos) {
        const PkLane &L = T.lanes[fin[pos]];
        int st = 0;
        for (int k = 0; k < CB_S_MAXSTEPS_ANY; ++k) el_of[pos][k] = 255;
        for (int q = 0; q < L.np; ++q) {
            const PkPart pt = L.parts[q];
            const CbPair p = in.pairs[T.blk[pt.blk]];
            for (int c = 0; c < pt.cnt; ++c, ++st) {
                const int32_t e = in.contribs[p.cstart + pt.c0 + c].e;
                int idx = 0;
                while (T.el[idx] != e) ++idx;
                el_of[pos][st] = (unsigned char)idx;
            }
        }
    }
    unsigned short clash_cnt[CB_PK_MAXEL][8];         // conflicts of shell e with colour c so far
    unsigned short degree[CB_PK_MAXEL];
    for (int e = 0; e < T.nel; ++e) { degree[e] = 0; for (int c = 0; c < 8; ++c) clash_cnt[e][c] = 0; }
    for (int o = 0; o < T.nlanes; o += 8)
        for (int st = 0; st < nsteps; ++st)
            for (int x = o; x < o + 8 && x < T.nlanes; ++x)
                for (int y = x + 1; y < o + 8 && y < T.nlanes; ++y)
                    if (el_of[x][st] != 255 && el_of[y][st] != 255 && el_of[x][st] != el_of[y][st]) {
                        ++degree[el_of[x][st]]; ++degree[el_of[y][st]];
                    }
    int eord[CB_PK_MAXEL];
    for (int e = 0; e < T.nel; ++e) eord[e] = e;
    for (int i = 1; i < T.nel; ++i) {                 // stable insertion sort, decreasing degree
        const int k = eord[i]; int j = i - 1;
        while (j >= 0 && degree[eord[j]] < degree[k]) { eord[j + 1] = eord[j]; --j; }
        eord[j + 1] = k;
    }
    signed char colour[CB_PK_MAXEL]; int cap[8], cnt8[8];
    for (int c = 0; c < 8; ++c) { cap[c] = 0; cnt8[c] = 0; }
    for (int k = 0; k < SL; ++k) ++cap[k & 7];
    for (int e = 0; e < T.nel; ++e) colour[e] = -1;
    for (int i = 0; i < T.nel; ++i) {
        const int e = eord[i];
        // conflicts with already coloured shells, per colour
        int cl[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int o = 0; o < T.nlanes; o += 8)
            for (int st = 0; st < nsteps; ++st) {
                bool mine = false;
                for (int x = o; x < o + 8 && x < T.nlanes; ++x) if (el_of[x][st] == e) mine = true;
                if (!mine) continue;
                for (int x = o; x < o + 8 && x < T.nlanes; ++x) {
                    const int f = el_of[x][st];
                    if (f != 255 && f != e && colour[f] >= 0) ++cl[(int)colour[f]];
                }
            }
        int best = -1;
        for (int c = 0; c < 8; ++c)
            if (cnt8[c] < cap[c] && (best < 0 || cl[c] < cl[best])) best = c;
        colour[e] = (signed char)best; ++cnt8[best];
    }
    (void)clash_cnt;
    int slot_of[CB_PK_MAXEL], nextslot[8] = {0, 1, 2, 3, 4, 5, 6, 7};
    for (int e = 0; e < T.nel; ++e) { slot_of[e] = nextslot[(int)colour[e]]; nextslot[(int)colour[e]] += 8; }
    // ---- records ----
    CbTileS t;
    t.out0 = T.out0; t.nout = T.nout; t.nsteps = (uint8_t)nsteps; t.np = (uint8_t)T.nblk; t.ne = (uint8_t)SL; t.pad = 0;
    out.tiles[ti] = t;
    uint32_t *steps = out.steps + ti * (long)S * 32;
    for (int i = 0; i < S * 32; ++i) steps[i] = CB_S_IDLE;
    for (int pos = 0; pos < T.nlanes; ++pos) {
        const PkLane &L = T.lanes[fin[pos]];
        int st = 0;
        for (int q = 0; q < L.np; ++q) {
            const PkPart pt = L.parts[q];
            const CbPair p = in.pairs[T.blk[pt.blk]];
            for (int c = 0; c < pt.cnt; ++c, ++st) {
                const CbContrib ct = in.contribs[p.cstart + pt.c0 + c];
                const int cls = in.cls ? in.cls[ct.e] : 0;
                steps[st * 32 + pos] = CB_S_REC(slot_of[el_of[pos][st]], ct.a, ct.b, c == 0, c == pt.cnt - 1, pt.follow, pt.blk, cls);
            }
        }
    }
    uint32_t *pr = out.pairsS + ti * (long)NP;
    for (int k = 0; k < NP; ++k) {
        if (k < T.nblk) {
            const CbPair p = in.pairs[T.blk[k]];
            pr[k] = CB_S_PAIR(p.off - T.out0, p.colh, p.maskA, p.maskB);
        } else pr[k] = 0;
        if (out.blk) out.blk[ti * (long)NP + k] = k < T.nblk ? T.blk[k] : -1;
    }
    int32_t *el = out.elems + ti * (long)SL;
    for (int k = 0; k < SL; ++k) el[k] = T.el[0];     // unused slots repeat a valid shell
    for (int e = 0; e < T.nel; ++e) el[slot_of[e]] = T.el[e];
}

// pack the joints [jbeg, jend); emits tiles ti0, ti0 + 1, ... when out != nullptr; returns the number of
// tiles, or -1 if the model does not fit the kernel shape (the caller falls back to another plan)
CB_HD inline long pk_segment(const PkIn &in, long jbeg, long jend, long ti0, const PkOut *out)
{
    const int S = in.shp.steps;
    PkTile T; T.nblk = 0; T.nel = 0; T.nlanes = 0; T.nout = 0; T.out0 = 0;
    bool open = false;
    long nt = 0;
    for (long B = jbeg; B < jend; ++B) {
        const int p0 = in.jpair[B], p1 = in.jpair[B + 1];
        if (p1 == p0 || !in.nfree[B]) continue;
        int cmax = 0;
        for (int q = p0; q < p1; ++q) if (in.pairs[q].ccount > cmax) cmax = in.pairs[q].ccount;
        const long out_n = (long)in.nfree[B] * in.colh[B];
        if (cmax > 2 * S || in.colh[B] > 255 || p1 - p0 > in.shp.pairs || p1 - p0 > 16 || out_n > in.shp.img) return -1;
        for (int attempt = 0; attempt < 2; ++attempt) {
            // shells of this joint that the tile does not hold yet
            int32_t newel[CB_PK_MAXEL]; int nnew = 0; bool too_many = false;
            for (int q = p0; q < p1 && !too_many; ++q)
                for (int c = 0; c < in.pairs[q].ccount; ++c) {
                    const int32_t e = in.contribs[in.pairs[q].cstart + c].e;
                    bool have = false;
                    for (int k = 0; k < T.nel && !have; ++k) have = T.el[k] == e;
                    for (int k = 0; k < nnew && !have; ++k) have = newel[k] == e;
                    if (!have) { if (nnew == CB_PK_MAXEL) { too_many = true; break; } newel[nnew++] = e; }
                }
            bool placed = false;
            if (!too_many && (!open || (in.base[B] - in.ax_base == T.out0 + T.nout)) && T.nout + out_n <= in.shp.img &&
                T.nel + nnew <= in.shp.slots && T.nblk + (p1 - p0) <= in.shp.pairs) {
                PkLane trial[32]; int ntrial = T.nlanes;
                for (int l = 0; l < T.nlanes; ++l) trial[l] = T.lanes[l];
                if (pk_place(in, trial, ntrial, p0, p1, T.nblk)) {
                    for (int l = 0; l < ntrial; ++l) T.lanes[l] = trial[l];
                    T.nlanes = ntrial; placed = true;
                }
            }
            if (placed) {
                if (!open) { T.out0 = in.base[B] - in.ax_base; open = true; }
                for (int k = 0; k < nnew; ++k) T.el[T.nel++] = newel[k];
                for (int q = p0; q < p1; ++q) T.blk[T.nblk++] = q;
                T.nout += (int32_t)out_n;
                break;
            }
            if (!open) return -1;                     // does not even fit an empty tile
            if (out) pk_close(in, T, ti0 + nt, *out);
            ++nt;
            T.nblk = 0; T.nel = 0; T.nlanes = 0; T.nout = 0; open = false;
        }
    }
    if (open) { if (out) pk_close(in, T, ti0 + nt, *out); ++nt; }
    return nt;
}
#endif
