/*
 * xtc_two_phase.c -- CPU prototype (test infrastructure, like the rest of oracle/) of the DEVICE decoder planned for
 * GROMACS XTC frames ("XTC decompression on device").  Not part of the product; the product's host decoder
 * is complexmixtures.jl_b200/csrc/cmx_xtc.inl, and tests/test_xtc_two_phase.py checks this restructuring against it.
 *
 * The compressed coordinate block is one serial bit stream: a "group" is
 *     [first atom: 3 integers in full range, `bitsize` bits] [flag: 1 bit] [if flag: 5-bit code]
 *     [run/3 following atoms: 3 small differences each, `smallidx` bits per atom]
 * with  code = run + is_smaller + 1  (run in {0,3,..,24}, is_smaller in {-1,0,+1}); without a flag the previous run
 * persists and is_smaller = 0; smallidx += is_smaller after the group.  Where a group starts therefore depends on all
 * groups before it -- but ONLY through the flag/code fields.  Hence two phases:
 *
 *   phase 1 (serial, cheap: touches 1 or 6 bits per group; the HOST reader thread does it):
 *       walk the stream and emit one byte per group, (flag << 7) | code;
 *   phase 2 (parallel; the DEVICE does it): from the group bytes, two prefix sums give every group's smallidx,
 *       first atom and bit offset; each group is then decoded independently (mixed-radix unpack of the first atom,
 *       the small differences chained inside the group, the swap of the first two atoms of a run).
 *
 * Shipped to the device per frame: the compressed block (~3.6 B/atom) + 1 B per group (~0.35 B/atom for water)
 * instead of 12 B/atom of decoded coordinates.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FIRSTIDX 9
static const int magicints[] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512,
                                645, 812, 1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007,
                                32768, 41285, 52015, 65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127, 524287, 660561,
                                832255, 1048576, 1321122, 1664510, 2097152, 2642245, 3329021, 4194304, 5284491, 6658042, 8388607,
                                10568983, 13316085, 16777216};
#define NMAGIC ((int)(sizeof(magicints) / sizeof(int)))

static uint32_t be32(const unsigned char *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

/* nbits (<= 32) of a big-endian bit stream starting at absolute bit `pos` -- random access, what a device thread does */
static uint32_t bits_at(const unsigned char *buf, size_t nbytes, uint64_t pos, int nbits) {
    uint64_t acc = 0;
    size_t b0 = (size_t)(pos >> 3);
    for (int k = 0; k < 8; ++k) acc = (acc << 8) | (b0 + (size_t)k < nbytes ? buf[b0 + k] : 0);
    int shift = 64 - (int)(pos & 7) - nbits;
    return (uint32_t)((acc >> shift) & (nbits >= 32 ? 0xffffffffull : ((1ull << nbits) - 1ull)));
}

/* three integers packed as one mixed-radix number of nbits bits, least-significant byte first (see cmx_xtc.inl) */
static void ints3_at(const unsigned char *buf, size_t nbytes, uint64_t pos, int nbits, const unsigned sizes[3], int out[3]) {
    unsigned bytes[32];
    int nb = 0, left = nbits;
    memset(bytes, 0, sizeof bytes);
    while (left > 8) { bytes[nb++] = bits_at(buf, nbytes, pos, 8); pos += 8; left -= 8; }
    if (left > 0) { bytes[nb++] = bits_at(buf, nbytes, pos, left); }
    for (int i = 2; i > 0; --i) {
        unsigned num = 0;
        for (int j = nb - 1; j >= 0; --j) {
            num = (num << 8) | bytes[j];
            unsigned p = num / sizes[i];
            bytes[j] = p;
            num -= p * sizes[i];
        }
        out[i] = (int)num;
    }
    out[0] = (int)(bytes[0] | (bytes[1] << 8) | (bytes[2] << 16) | (bytes[3] << 24));
}

static int sizeofint(unsigned size) { unsigned num = 1; int n = 0; while (size >= num && n < 32) { n++; num <<= 1; } return n; }
static int sizeofints(const unsigned s[3]) {
    unsigned bytes[32]; int nbytes = 1; bytes[0] = 1;
    for (int i = 0; i < 3; ++i) {
        unsigned tmp = 0; int k = 0;
        for (; k < nbytes; ++k) { tmp = bytes[k] * s[i] + tmp; bytes[k] = tmp & 0xff; tmp >>= 8; }
        while (tmp) { bytes[k++] = tmp & 0xff; tmp >>= 8; }
        nbytes = k;
    }
    unsigned num = 1; int n = 0; nbytes--;
    while (bytes[nbytes] >= num) { n++; num *= 2; }
    return n + nbytes * 8;
}

typedef struct {
    float precision; int minint[3]; unsigned sizeint[3]; int bitsizeint[3]; int bitsize, smallidx0; uint32_t nbytes;
    const unsigned char *stream;
} xtc_hdr;

static int parse_hdr(const unsigned char *p, size_t avail, xtc_hdr *h) {
    if (avail < 36) return 0;
    uint32_t u = be32(p); memcpy(&h->precision, &u, 4);
    int maxint[3];
    for (int k = 0; k < 3; ++k) { h->minint[k] = (int)be32(p + 4 + 4 * k); maxint[k] = (int)be32(p + 16 + 4 * k); h->sizeint[k] = (unsigned)(maxint[k] - h->minint[k]) + 1u; }
    h->smallidx0 = (int)be32(p + 28); h->nbytes = be32(p + 32); h->stream = p + 36;
    if ((size_t)h->nbytes + 36 > avail || h->smallidx0 < FIRSTIDX || h->smallidx0 >= NMAGIC) return 0;
    if ((h->sizeint[0] | h->sizeint[1] | h->sizeint[2]) > 0xffffffu) {
        for (int k = 0; k < 3; ++k) h->bitsizeint[k] = sizeofint(h->sizeint[k]);
        h->bitsize = 0;
    } else { h->bitsize = sizeofints(h->sizeint); h->bitsizeint[0] = h->bitsizeint[1] = h->bitsizeint[2] = 0; }
    return 1;
}
static int first_bits(const xtc_hdr *h) { return h->bitsize ? h->bitsize : h->bitsizeint[0] + h->bitsizeint[1] + h->bitsizeint[2]; }

/* ---- phase 1: the serial walk.  block = coordinate block of a frame of natoms > 9 atoms (starts at `precision`).
 * codes[g] = (flag << 7) | code.  Returns the number of groups (0 = malformed / codes too small). */
int xtc_skeleton(const unsigned char *block, size_t avail, int natoms, unsigned char *codes, int max_groups) {
    xtc_hdr h;
    if (!parse_hdr(block, avail, &h)) return 0;
    const int fb = first_bits(&h);
    uint64_t pos = 0;
    int i = 0, run = 0, smallidx = h.smallidx0, g = 0;
    while (i < natoms) {
        if (g >= max_groups) return 0;
        pos += (uint64_t)fb;
        unsigned flag = bits_at(h.stream, h.nbytes, pos, 1); pos += 1;
        unsigned code = 0;
        int is_smaller = 0;
        if (flag) {
            code = bits_at(h.stream, h.nbytes, pos, 5); pos += 5;
            run = (int)code; is_smaller = run % 3; run -= is_smaller; is_smaller--;
        }
        codes[g++] = (unsigned char)((flag << 7) | code);
        pos += (uint64_t)(run / 3) * (uint64_t)smallidx;
        i += 1 + run / 3;
        smallidx += is_smaller;
        if (smallidx < FIRSTIDX || smallidx >= NMAGIC || (pos + 7) / 8 > (uint64_t)h.nbytes + 8) return 0;
    }
    return i == natoms ? g : 0;
}

/* ---- phase 2: what the device does.  The three "scans" are written as loops; every iteration of the final loop is
 * independent of the others (one thread per group). */
int xtc_decode_from_skeleton(const unsigned char *block, size_t avail, int natoms, const unsigned char *codes, int ngroups,
                             float *xyz /* [natoms][3] Angstrom */) {
    xtc_hdr h;
    if (!parse_hdr(block, avail, &h)) return 0;
    const int fb = first_bits(&h);
    int *run = (int *)malloc(sizeof(int) * (size_t)ngroups), *sidx = (int *)malloc(sizeof(int) * (size_t)ngroups);
    int *atom0 = (int *)malloc(sizeof(int) * (size_t)ngroups);
    uint64_t *bitpos = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)ngroups);
    /* scan 1: run[g] = run of the last flagged group at or before g ("last writer wins": a max-scan over flagged
     * positions); smallidx[g] = smallidx0 + sum of is_smaller over groups before g (a plus-scan) */
    int cur_run = 0, cur_sidx = h.smallidx0;
    for (int g = 0; g < ngroups; ++g) {
        int is_smaller = 0;
        if (codes[g] & 0x80) { int c = codes[g] & 31; is_smaller = c % 3; cur_run = c - is_smaller; is_smaller--; }
        run[g] = cur_run; sidx[g] = cur_sidx;
        cur_sidx += is_smaller;
    }
    /* scan 2: first atom and bit offset of every group (plus-scans of per-group sizes) */
    int a = 0; uint64_t pos = 0;
    for (int g = 0; g < ngroups; ++g) {
        atom0[g] = a; bitpos[g] = pos;
        a += 1 + run[g] / 3;
        pos += (uint64_t)fb + 1 + ((codes[g] & 0x80) ? 5 : 0) + (uint64_t)(run[g] / 3) * (uint64_t)sidx[g];
    }
    int ok = a == natoms;
    const float inv_precision = 1.0f / h.precision;
    /* the parallel part: one independent task per group */
    for (int g = 0; g < ngroups && ok; ++g) {
        uint64_t p = bitpos[g];
        int cur[3], prev[3];
        if (h.bitsize == 0) {
            for (int k = 0; k < 3; ++k) { cur[k] = (int)bits_at(h.stream, h.nbytes, p, h.bitsizeint[k]); p += (uint64_t)h.bitsizeint[k]; }
        } else { ints3_at(h.stream, h.nbytes, p, h.bitsize, h.sizeint, cur); p += (uint64_t)h.bitsize; }
        p += 1 + ((codes[g] & 0x80) ? 5 : 0);
        for (int k = 0; k < 3; ++k) { cur[k] += h.minint[k]; prev[k] = cur[k]; }
        float *out = xyz + 3 * (size_t)atom0[g];
        const int nsmall = run[g] / 3;
        if (nsmall == 0) { for (int k = 0; k < 3; ++k) out[k] = (float)((double)((float)cur[k] * inv_precision) * 10.0); continue; }
        const int smallnum = magicints[sidx[g]] / 2;
        const unsigned ss[3] = {(unsigned)magicints[sidx[g]], (unsigned)magicints[sidx[g]], (unsigned)magicints[sidx[g]]};
        for (int q = 0; q < nsmall; ++q) {
            int nxt[3];
            ints3_at(h.stream, h.nbytes, p, sidx[g], ss, nxt); p += (uint64_t)sidx[g];
            for (int k = 0; k < 3; ++k) nxt[k] += prev[k] - smallnum;
            if (q == 0) {
                for (int k = 0; k < 3; ++k) { int t = nxt[k]; nxt[k] = prev[k]; prev[k] = t; }
                for (int k = 0; k < 3; ++k) *out++ = (float)((double)((float)prev[k] * inv_precision) * 10.0);
            } else {
                for (int k = 0; k < 3; ++k) prev[k] = nxt[k];
            }
            for (int k = 0; k < 3; ++k) *out++ = (float)((double)((float)nxt[k] * inv_precision) * 10.0);
        }
    }
    free(run); free(sidx); free(atom0); free(bitpos);
    return ok;
}
