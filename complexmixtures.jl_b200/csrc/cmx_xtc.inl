// cmx_xtc.inl -- native GROMACS XTC reader (host code, included by cmx_b200.cu): SURVEY 8 (f1), second format.
//
// The reference reads XTC through Chemfiles (src/trajectory_formats/ChemFiles.jl:112-138: read_step, positions in
// Angstrom, unit cell matrix); this is a from-the-format-description decoder of the XTC frame layout and of its
// compressed coordinate block (the "xdr3dfcoord" scheme: coordinates quantised to integers at `precision` per nm, the
// first atom of a run stored in full range with a mixed-radix packing of the three components, the following atoms
// as small differences in an adaptive range; the first two atoms of a run are swapped, which favours water).
// Output: fp32 xyz triplets in Angstrom (nm x 10 in double, rounded once to fp32) and the cell as the column-major 3x3
// matrix cmx_submit_frame takes.  Frames are indexed at open time so that any frame can be read by number and ranks can
// read disjoint frames.
//
// Two decoders share the format description above.  The HOST decoder (xtc_decode_coords) serves cmx_xtc_read_frame and
// the frames of at most 9 atoms (stored as plain floats).  The DEVICE decoder serves the native feed (cmx_run_xtc): the
// compressed block is one serial bit stream, but where a group of atoms starts depends on the earlier groups only
// through their 1-bit flag / 5-bit run code, so the reader thread walks just those fields (xtc_skeleton: a few bits
// per group) and emits one 12-byte record per group -- first atom, bit offset, small-difference range, run length --
// and the device decodes every group with its own thread (k_xtc_decode: mixed-radix unpack of the first atom, the
// chained small differences, the swap of the first two atoms of a run).  The host ships the compressed block and the
// records (~ 4 + 4.5 B/atom for water) instead of decoding 12 B/atom at ~1.3 ms per 100 k atoms and thread.

struct cmx_xtc {
    int fd = -1;
    cmx_xtc_info info{};
    std::vector<int64_t> offset;   // byte offset of every frame
    std::vector<int64_t> length;   // bytes of every frame
    std::string path;
};

namespace {

// first index of the table below that holds a usable range
constexpr int XTC_FIRSTIDX = 9;
// "magic" ranges of the adaptive small-difference coder: roughly 2^(k/3)
const int xtc_magicints[] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512,
                             645, 812, 1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007,
                             32768, 41285, 52015, 65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127, 524287, 660561,
                             832255, 1048576, 1321122, 1664510, 2097152, 2642245, 3329021, 4194304, 5284491, 6658042, 8388607,
                             10568983, 13316085, 16777216};
constexpr int XTC_NMAGIC = (int)(sizeof(xtc_magicints) / sizeof(int));

inline uint32_t be32(const unsigned char *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3]; }
inline int32_t be32i(const unsigned char *p) { return (int32_t)be32(p); }
inline float be32f(const unsigned char *p) { uint32_t u = be32(p); float f; std::memcpy(&f, &u, 4); return f; }

struct XtcBits {   // big-endian bit reader over the compressed block: 64-bit window refilled bytewise
    const unsigned char *buf; size_t n, cnt = 0; uint64_t acc = 0; int nacc = 0; bool overrun = false;
    inline void refill(int need) {
        while (nacc < need) {
            uint64_t b = 0;
            if (cnt < n) b = buf[cnt++]; else overrun = true;
            acc = (acc << 8) | b; nacc += 8;
        }
    }
    // nbits <= 32: the next nbits of the stream, most significant bit first
    inline uint32_t bits(int nbits) {
        if (nbits == 0) return 0;
        refill(nbits);
        nacc -= nbits;
        return (uint32_t)((acc >> nacc) & ((nbits >= 32) ? 0xffffffffull : ((1ull << nbits) - 1ull)));
    }
    // Three integers packed as one mixed-radix number of `nbits` bits with radices sizes[0..2].  The number is stored
    // least-significant BYTE first (whole bytes, then the remaining high bits in one piece).
    void ints3(int nbits, const unsigned sizes[3], int out[3]) {
        if (nbits <= 64) {   // the usual case: assemble the number in a machine word, two divisions
            uint64_t v = 0; int shift = 0, left = nbits;
            while (left > 8) { v |= (uint64_t)bits(8) << shift; shift += 8; left -= 8; }
            if (left > 0) v |= (uint64_t)bits(left) << shift;
            const uint64_t q2 = v / sizes[2];
            out[2] = (int)(v - q2 * sizes[2]);
            const uint64_t q1 = q2 / sizes[1];
            out[1] = (int)(q2 - q1 * sizes[1]);
            out[0] = (int)q1;
            return;
        }
        unsigned bytes[32] = {0};
        int nb = 0;
        while (nbits > 8) { bytes[nb++] = bits(8); nbits -= 8; }
        if (nbits > 0) bytes[nb++] = bits(nbits);
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
};

int xtc_sizeofint(unsigned size) {
    unsigned num = 1; int nbits = 0;
    while (size >= num && nbits < 32) { nbits++; num <<= 1; }
    return nbits;
}
// bits needed for the product of three ranges (byte-wise multi-precision product)
int xtc_sizeofints(const unsigned sizes[3]) {
    unsigned bytes[32]; int nbytes = 1; bytes[0] = 1;
    for (int i = 0; i < 3; ++i) {
        unsigned tmp = 0; int k = 0;
        for (; k < nbytes; ++k) { tmp = bytes[k] * sizes[i] + tmp; bytes[k] = tmp & 0xff; tmp >>= 8; }
        while (tmp != 0) { bytes[k++] = tmp & 0xff; tmp >>= 8; }
        nbytes = k;
    }
    unsigned num = 1; int nbits = 0;
    nbytes--;
    while (bytes[nbytes] >= num) { nbits++; num *= 2; }
    return nbits + nbytes * 8;
}

// frame header: magic(1995) natoms step time | box[9] | natoms ; returns header bytes or 0
constexpr int XTC_HEADER = 4 * 4 + 9 * 4 + 4;

// Decodes the coordinate block that follows the frame header.  `p` points at it, `avail` bytes are readable.
// Returns the number of bytes consumed (0 on a malformed block).
size_t xtc_decode_coords(const unsigned char *p, size_t avail, int natoms, float *xyz_out /* [natoms][3], Angstrom */) {
    if (natoms <= 9) {   // small systems are stored as plain floats
        size_t need = (size_t)natoms * 12;
        if (avail < need) return 0;
        for (int k = 0; k < 3 * natoms; ++k) xyz_out[k] = (float)((double)be32f(p + 4 * k) * 10.0);
        return need;
    }
    if (avail < 36) return 0;
    const float precision = be32f(p);
    int minint[3], maxint[3];
    for (int k = 0; k < 3; ++k) { minint[k] = be32i(p + 4 + 4 * k); maxint[k] = be32i(p + 16 + 4 * k); }
    int smallidx = be32i(p + 28);
    const uint32_t nbytes = be32(p + 32);
    if (!(precision > 0) || smallidx < XTC_FIRSTIDX || smallidx >= XTC_NMAGIC || (size_t)nbytes + 36 > avail) return 0;
    unsigned sizeint[3], bitsizeint[3] = {0, 0, 0};
    for (int k = 0; k < 3; ++k) {
        if (maxint[k] < minint[k]) return 0;
        sizeint[k] = (unsigned)(maxint[k] - minint[k]) + 1u;
    }
    int bitsize;
    if ((sizeint[0] | sizeint[1] | sizeint[2]) > 0xffffffu) {
        for (int k = 0; k < 3; ++k) bitsizeint[k] = (unsigned)xtc_sizeofint(sizeint[k]);
        bitsize = 0;   // the three components are stored separately
    } else bitsize = xtc_sizeofints(sizeint);
    int smaller = xtc_magicints[std::max(XTC_FIRSTIDX, smallidx - 1)] / 2;
    int smallnum = xtc_magicints[smallidx] / 2;
    unsigned sizesmall[3] = {(unsigned)xtc_magicints[smallidx], (unsigned)xtc_magicints[smallidx], (unsigned)xtc_magicints[smallidx]};
    // the format defines a coordinate as float(int) * float(1/precision) nm; reproduce that value, then scale to Angstrom
    const float inv_precision = 1.0f / precision;
    XtcBits br{p + 36, nbytes};
    float *out = xyz_out;
    auto emit = [&](const int c[3]) {
        for (int k = 0; k < 3; ++k) *out++ = (float)((double)((float)c[k] * inv_precision) * 10.0);
    };
    int i = 0, run = 0;
    while (i < natoms) {
        int cur[3], prev[3];
        if (bitsize == 0) { for (int k = 0; k < 3; ++k) cur[k] = (int)br.bits((int)bitsizeint[k]); }
        else br.ints3(bitsize, sizeint, cur);
        ++i;
        for (int k = 0; k < 3; ++k) { cur[k] += minint[k]; prev[k] = cur[k]; }
        const uint32_t flag = br.bits(1);
        int is_smaller = 0;
        if (flag == 1) {
            run = (int)br.bits(5);
            is_smaller = run % 3;
            run -= is_smaller;
            is_smaller--;
        }
        if (run > 0) {
            if (i + run / 3 > natoms) return 0;
            for (int k = 0; k < run; k += 3) {
                int nxt[3];
                br.ints3(smallidx, sizesmall, nxt);
                ++i;
                for (int q = 0; q < 3; ++q) nxt[q] += prev[q] - smallnum;
                if (k == 0) {
                    // the first two atoms of a run are stored swapped (oxygen after the first hydrogen of a water)
                    for (int q = 0; q < 3; ++q) std::swap(nxt[q], prev[q]);
                    emit(prev);
                } else {
                    for (int q = 0; q < 3; ++q) prev[q] = nxt[q];
                }
                emit(nxt);
            }
        } else emit(cur);
        smallidx += is_smaller;
        if (smallidx < XTC_FIRSTIDX || smallidx >= XTC_NMAGIC) return 0;
        if (is_smaller < 0) {
            smallnum = smaller;
            smaller = smallidx > XTC_FIRSTIDX ? xtc_magicints[smallidx - 1] / 2 : 0;
        } else if (is_smaller > 0) {
            smaller = smallnum;
            smallnum = xtc_magicints[smallidx] / 2;
        }
        sizesmall[0] = sizesmall[1] = sizesmall[2] = (unsigned)xtc_magicints[smallidx];
        if (br.overrun) return 0;
    }
    if (out != xyz_out + 3 * (size_t)natoms) return 0;
    return 36 + (((size_t)nbytes + 3) & ~(size_t)3);
}

// bytes of the frame that starts at `off` (header + coordinate block), 0 at end of file / malformed
int64_t xtc_frame_length(int fd, int64_t off, int64_t file_size, int *natoms_out) {
    unsigned char h[XTC_HEADER + 36];
    if (off + XTC_HEADER > file_size) return 0;
    size_t want = (size_t)std::min<int64_t>((int64_t)sizeof h, file_size - off);
    if (!pread_all(fd, h, want, (off_t)off)) return 0;
    if (be32i(h) != 1995) return 0;
    const int natoms = be32i(h + 4);
    if (natoms <= 0 || be32i(h + XTC_HEADER - 4) != natoms) return 0;
    *natoms_out = natoms;
    if (natoms <= 9) return XTC_HEADER + 12 * (int64_t)natoms;
    if (want < sizeof h) return 0;
    const uint32_t nbytes = be32(h + XTC_HEADER + 32);
    return XTC_HEADER + 36 + (int64_t)((nbytes + 3u) & ~3u);
}

// ---- device decoder ------------------------------------------------------------------------------------------------
// slot layout: [XtcDevHeader (64 B)][XtcGroup records][compressed bit stream + 16 zero bytes]
struct XtcDevHeader {
    int32_t natoms, ngroups;
    float inv_precision;
    int32_t minint[3];
    uint32_t sizeint[3];
    int32_t bitsizeint[3];
    int32_t bitsize;
    uint32_t nbytes, rec_off, stream_off;
};
static_assert(sizeof(XtcDevHeader) == 64, "XtcDevHeader is the first 64 bytes of the slot");
struct XtcGroup {
    uint32_t atom0;        // first atom of the group
    uint32_t bitpos;       // bit offset of the group in the stream
    uint8_t sidx;          // index of the small-difference range in force for this group
    uint8_t nsmall;        // atoms that follow the first one (run / 3)
    uint8_t ctl_bits;      // 1 (flag only) or 6 (flag + 5-bit code)
    uint8_t pad;
};
static_assert(sizeof(XtcGroup) == 12, "XtcGroup is 12 bytes");

__constant__ int c_xtc_magic[XTC_NMAGIC];

// nbits (<= 32) of the big-endian bit stream at absolute bit `pos` (random access)
__device__ __forceinline__ uint32_t xtc_bits_at(const unsigned char *__restrict__ buf, uint32_t pos, int nbits) {
    const uint32_t b0 = pos >> 3;
    unsigned long long acc = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc = (acc << 8) | (unsigned long long)buf[b0 + k];      // (the stream is padded with 16 zero bytes)
    const int shift = 64 - (int)(pos & 7u) - nbits;
    return (uint32_t)((acc >> shift) & (nbits >= 32 ? 0xffffffffull : ((1ull << nbits) - 1ull)));
}
// three integers packed as one mixed-radix number of nbits bits, least-significant BYTE first
__device__ __forceinline__ void xtc_ints3_at(const unsigned char *__restrict__ buf, uint32_t pos, int nbits, const uint32_t sizes[3], int out[3]) {
    if (nbits <= 64) {
        unsigned long long v = 0; int shift = 0, left = nbits;
        while (left > 8) { v |= (unsigned long long)xtc_bits_at(buf, pos, 8) << shift; pos += 8; shift += 8; left -= 8; }
        if (left > 0) v |= (unsigned long long)xtc_bits_at(buf, pos, left) << shift;
        const unsigned long long q2 = v / sizes[2];
        out[2] = (int)(v - q2 * sizes[2]);
        const unsigned long long q1 = q2 / sizes[1];
        out[1] = (int)(q2 - q1 * sizes[1]);
        out[0] = (int)q1;
        return;
    }
    uint32_t bytes[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};      // nbits <= 3 * 24 = 72 when the packed form is used
    int nb = 0, left = nbits;
    while (left > 8) { bytes[nb++] = xtc_bits_at(buf, pos, 8); pos += 8; left -= 8; }
    if (left > 0) bytes[nb++] = xtc_bits_at(buf, pos, left);
    for (int i = 2; i > 0; --i) {
        uint32_t num = 0;
        for (int j = nb - 1; j >= 0; --j) {
            num = (num << 8) | bytes[j];
            const uint32_t p = num / sizes[i];
            bytes[j] = p;
            num -= p * sizes[i];
        }
        out[i] = (int)num;
    }
    out[0] = (int)(bytes[0] | (bytes[1] << 8) | (bytes[2] << 16) | (bytes[3] << 24));
}
// the value the format defines -- float(int) * float(1/precision) nm -- scaled to Angstrom like the host decoder does
__device__ __forceinline__ float xtc_coord(int c, float inv_precision) {
    return (float)__dmul_rn((double)__fmul_rn((float)c, inv_precision), 10.0);
}

__global__ void __launch_bounds__(128)
k_xtc_decode(const unsigned char *__restrict__ raw, float *__restrict__ xyz) {
    const XtcDevHeader H = *reinterpret_cast<const XtcDevHeader *>(raw);
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= H.ngroups) return;
    const XtcGroup G = reinterpret_cast<const XtcGroup *>(raw + H.rec_off)[g];
    const unsigned char *stream = raw + H.stream_off;
    uint32_t p = G.bitpos;
    int cur[3], prev[3];
    if (H.bitsize == 0) {
        for (int k = 0; k < 3; ++k) { cur[k] = (int)xtc_bits_at(stream, p, H.bitsizeint[k]); p += (uint32_t)H.bitsizeint[k]; }
    } else { xtc_ints3_at(stream, p, H.bitsize, H.sizeint, cur); p += (uint32_t)H.bitsize; }
    p += G.ctl_bits;
    for (int k = 0; k < 3; ++k) { cur[k] += H.minint[k]; prev[k] = cur[k]; }
    float *out = xyz + 3 * (size_t)G.atom0;
    if (G.nsmall == 0) { for (int k = 0; k < 3; ++k) out[k] = xtc_coord(cur[k], H.inv_precision); return; }
    const int magic = c_xtc_magic[G.sidx], smallnum = magic / 2;
    const uint32_t ss[3] = {(uint32_t)magic, (uint32_t)magic, (uint32_t)magic};
    for (int q = 0; q < (int)G.nsmall; ++q) {
        int nxt[3];
        xtc_ints3_at(stream, p, (int)G.sidx, ss, nxt); p += G.sidx;
        for (int k = 0; k < 3; ++k) nxt[k] += prev[k] - smallnum;
        if (q == 0) {      // the first two atoms of a run are stored swapped
            for (int k = 0; k < 3; ++k) { const int t = nxt[k]; nxt[k] = prev[k]; prev[k] = t; }
            for (int k = 0; k < 3; ++k) *out++ = xtc_coord(prev[k], H.inv_precision);
        } else {
            for (int k = 0; k < 3; ++k) prev[k] = nxt[k];
        }
        for (int k = 0; k < 3; ++k) *out++ = xtc_coord(nxt[k], H.inv_precision);
    }
}

std::once_flag g_xtc_magic_once[64];      // the table is uploaded once per device

// The reader thread's share of a compressed frame: parse the block header and walk the flag / run-code fields of the
// bit stream (1 or 6 bits per group), emitting the group records.  `block` points at the coordinate block (after the
// frame header), `slot` receives [XtcDevHeader][records][stream + padding]; returns the bytes of the slot to copy (0 =
// malformed / slot too small) and the number of groups.
size_t xtc_skeleton(const unsigned char *block, size_t avail, int natoms, unsigned char *slot, size_t slot_bytes, int *ngroups_out) {
    if (avail < 36) return 0;
    XtcDevHeader H{};
    const float precision = be32f(block);
    int maxint[3];
    for (int k = 0; k < 3; ++k) {
        H.minint[k] = be32i(block + 4 + 4 * k); maxint[k] = be32i(block + 16 + 4 * k);
        if (maxint[k] < H.minint[k]) return 0;
        H.sizeint[k] = (uint32_t)(maxint[k] - H.minint[k]) + 1u;
    }
    int smallidx = be32i(block + 28);
    H.nbytes = be32(block + 32);
    if (!(precision > 0) || smallidx < XTC_FIRSTIDX || smallidx >= XTC_NMAGIC || (size_t)H.nbytes + 36 > avail) return 0;
    if ((H.sizeint[0] | H.sizeint[1] | H.sizeint[2]) > 0xffffffu) {
        for (int k = 0; k < 3; ++k) H.bitsizeint[k] = xtc_sizeofint(H.sizeint[k]);
        H.bitsize = 0;
    } else H.bitsize = xtc_sizeofints(H.sizeint);
    const int fb = H.bitsize ? H.bitsize : H.bitsizeint[0] + H.bitsizeint[1] + H.bitsizeint[2];
    H.natoms = natoms; H.inv_precision = 1.0f / precision;
    H.rec_off = (uint32_t)sizeof(XtcDevHeader);
    XtcGroup *rec = reinterpret_cast<XtcGroup *>(slot + H.rec_off);
    const size_t max_groups = (slot_bytes - sizeof(XtcDevHeader) - (size_t)H.nbytes - 32) / sizeof(XtcGroup);
    const unsigned char *stream = block + 36;
    uint64_t pos = 0;
    int i = 0, run = 0, ng = 0;
    auto bits_at = [&](uint64_t at, int nbits) -> uint32_t {
        uint64_t acc = 0; const size_t b0 = (size_t)(at >> 3);
        for (int k = 0; k < 8; ++k) acc = (acc << 8) | (b0 + (size_t)k < (size_t)H.nbytes ? stream[b0 + k] : 0);
        return (uint32_t)((acc >> (64 - (int)(at & 7) - nbits)) & ((1ull << nbits) - 1ull));
    };
    while (i < natoms) {
        if ((size_t)ng >= max_groups) return 0;
        XtcGroup G{};
        G.atom0 = (uint32_t)i; G.bitpos = (uint32_t)pos; G.sidx = (uint8_t)smallidx;
        pos += (uint64_t)fb;
        const uint32_t flag = bits_at(pos, 1); pos += 1;
        int is_smaller = 0;
        G.ctl_bits = 1;
        if (flag) {
            run = (int)bits_at(pos, 5); pos += 5;
            is_smaller = run % 3; run -= is_smaller; is_smaller--;
            G.ctl_bits = 6;
        }
        G.nsmall = (uint8_t)(run / 3);
        rec[ng++] = G;
        pos += (uint64_t)(run / 3) * (uint64_t)smallidx;
        i += 1 + run / 3;
        smallidx += is_smaller;
        if (smallidx < XTC_FIRSTIDX || smallidx >= XTC_NMAGIC || (pos + 7) / 8 > (uint64_t)H.nbytes + 8 || pos > 0xfffffff0ull) return 0;
    }
    if (i != natoms) return 0;
    H.ngroups = ng;
    H.stream_off = (uint32_t)((H.rec_off + sizeof(XtcGroup) * (size_t)ng + 15) & ~(size_t)15);
    std::memcpy(slot + H.stream_off, stream, H.nbytes);
    std::memset(slot + H.stream_off + H.nbytes, 0, 16);
    std::memcpy(slot, &H, sizeof H);
    *ngroups_out = ng;
    return (size_t)H.stream_off + H.nbytes + 16;
}

}  // namespace

void launch_xtc_decode(const unsigned char *d_raw, float *d_dec, int ngroups, cudaStream_t stream) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::call_once(g_xtc_magic_once[dev & 63], [] { cudaMemcpyToSymbol(c_xtc_magic, xtc_magicints, sizeof(int) * XTC_NMAGIC); });
    if (ngroups > 0) k_xtc_decode<<<(unsigned)((ngroups + 127) / 128), 128, 0, stream>>>(d_raw, d_dec);
}

extern "C" {

int32_t cmx_xtc_open(const char *path, cmx_xtc **out, cmx_xtc_info *info) {
    if (!path || !out) return dcd_fail(CMX_ERR_ARG, "cmx_xtc_open: null argument");
    *out = nullptr;
    int fd = open(path, O_RDONLY);
    if (fd < 0) return dcd_fail(CMX_ERR_IO, std::string("cannot open ") + path);
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return dcd_fail(CMX_ERR_IO, std::string("cannot stat ") + path); }
    cmx_xtc *x = new cmx_xtc();
    x->fd = fd; x->path = path;
    int64_t off = 0;
    int natoms0 = -1;
    while (off < (int64_t)st.st_size) {
        int natoms = 0;
        int64_t len = xtc_frame_length(fd, off, (int64_t)st.st_size, &natoms);
        if (len <= 0 || off + len > (int64_t)st.st_size) break;   // a truncated last frame is ignored
        if (natoms0 < 0) natoms0 = natoms;
        if (natoms != natoms0) { close(fd); delete x; return dcd_fail(CMX_ERR_IO, "XTC file with a varying number of atoms"); }
        x->offset.push_back(off); x->length.push_back(len);
        off += len;
    }
    if (x->offset.empty()) { close(fd); delete x; return dcd_fail(CMX_ERR_IO, "not an XTC file (magic 1995) or no complete frame"); }
    x->info.natoms = natoms0; x->info.nframes = (int64_t)x->offset.size();
    if (info) *info = x->info;
    *out = x;
    return CMX_OK;
}

int32_t cmx_xtc_close(cmx_xtc *x) {
    if (!x) return CMX_OK;
    if (x->fd >= 0) close(x->fd);
    delete x;
    return CMX_OK;
}

static bool xtc_read_into(const cmx_xtc *x, int64_t iframe, std::vector<unsigned char> &buf, float *xyz, double cell[9],
                          int32_t *step, float *time, std::string &err) {
    buf.resize((size_t)x->length[(size_t)iframe]);
    if (!pread_all(x->fd, buf.data(), buf.size(), (off_t)x->offset[(size_t)iframe])) { err = "short read (XTC frame)"; return false; }
    const unsigned char *h = buf.data();
    if (step) *step = be32i(h + 8);
    if (time) *time = be32f(h + 12);
    if (cell) {
        // box vectors are the ROWS of the stored 3x3 (nm); cmx takes the lattice vectors as columns, column-major:
        // cell[3*v + k] = component k of vector v, in Angstrom
        for (int v = 0; v < 3; ++v)
            for (int k = 0; k < 3; ++k) cell[3 * v + k] = (double)be32f(h + 16 + 4 * (3 * v + k)) * 10.0;
    }
    if (xyz && xtc_decode_coords(h + XTC_HEADER, buf.size() - XTC_HEADER, (int)x->info.natoms, xyz) == 0) {
        err = "malformed XTC coordinate block in frame " + std::to_string((long long)iframe);
        return false;
    }
    return true;
}

int32_t cmx_xtc_read_frame(cmx_xtc *x, int64_t iframe, float *xyz, double cell[9], int32_t *step, float *time) {
    if (!x) return dcd_fail(CMX_ERR_ARG, "cmx_xtc_read_frame: null handle");
    if (iframe < 0 || iframe >= x->info.nframes) return dcd_fail(CMX_ERR_ARG, "cmx_xtc_read_frame: frame out of range");
    std::vector<unsigned char> buf;
    std::string err;
    if (!xtc_read_into(x, iframe, buf, xyz, cell, step, time, err)) return dcd_fail(CMX_ERR_IO, err);
    return CMX_OK;
}

// One frame decoded ON THE DEVICE (the decoder of the native feed, callable by itself): the reader's skeleton walk on
// the host, k_xtc_decode on `device`, the coordinates copied back.  Bit-identical to cmx_xtc_read_frame; frames of at
// most 9 atoms (plain floats in the file) are returned by the host decoder.
int32_t cmx_xtc_read_frame_device(cmx_xtc *x, int64_t iframe, int32_t device, float *xyz, double cell[9]) {
    if (!x || !xyz) return dcd_fail(CMX_ERR_ARG, "cmx_xtc_read_frame_device: null argument");
    if (iframe < 0 || iframe >= x->info.nframes) return dcd_fail(CMX_ERR_ARG, "cmx_xtc_read_frame_device: frame out of range");
    if (x->info.natoms <= 9) return cmx_xtc_read_frame(x, iframe, xyz, cell, nullptr, nullptr);
    std::vector<unsigned char> buf;
    std::string err;
    if (!xtc_read_into(x, iframe, buf, nullptr, cell, nullptr, nullptr, err)) return dcd_fail(CMX_ERR_IO, err);
    const size_t natoms = (size_t)x->info.natoms;
    const size_t slot_bytes = sizeof(XtcDevHeader) + sizeof(XtcGroup) * natoms + buf.size() + 64;
    std::vector<unsigned char> slot(slot_bytes);
    int ng = 0;
    const size_t used = xtc_skeleton(buf.data() + XTC_HEADER, buf.size() - XTC_HEADER, (int)natoms, slot.data(), slot_bytes, &ng);
    if (!used) return dcd_fail(CMX_ERR_IO, "malformed XTC coordinate block in frame " + std::to_string((long long)iframe));
    auto cuda_fail = [](cudaError_t e) { return dcd_fail(CMX_ERR_CUDA, std::string("cmx_xtc_read_frame_device: ") + cudaGetErrorString(e)); };
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(e);
    unsigned char *d_raw = nullptr; float *d_dec = nullptr;
    if ((e = cudaMalloc(&d_raw, used)) != cudaSuccess) return cuda_fail(e);
    if ((e = cudaMalloc(&d_dec, 12 * natoms)) != cudaSuccess) { cudaFree(d_raw); return cuda_fail(e); }
    e = cudaMemcpy(d_raw, slot.data(), used, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) { launch_xtc_decode(d_raw, d_dec, ng, nullptr); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaMemcpy(xyz, d_dec, 12 * natoms, cudaMemcpyDeviceToHost);
    cudaFree(d_raw); cudaFree(d_dec);
    return e == cudaSuccess ? CMX_OK : cuda_fail(e);
}

// The frame loop for an XTC file: like cmx_run_dcd, but the reader threads also DECODE (the compressed block is a
// serial bit stream per frame, so frames are the unit of parallelism: one frame per thread at a time); the ring
// holds decoded fp32 xyz triplets of the whole frame, the selection gather runs on the device.
int32_t cmx_run_xtc(cmx_handle *h, cmx_xtc *x, const int32_t *solute_indices, const int32_t *solvent_indices,
                    const int64_t *frames, const double *weights, int64_t nframes, int32_t n_reader_threads) {
    if (!h) return CMX_ERR_ARG;
    if (!x) return fail(h, CMX_ERR_ARG, "cmx_run_xtc: null argument");
    FeedSource src;
    src.what = "cmx_run_xtc"; src.natoms = x->info.natoms; src.nframes = x->info.nframes;
    const bool on_device = x->info.natoms > 9 && !(is_group(h) ? h->children[0] : h)->xtc_host_decode;
    if (!on_device) {
        // frames of <= 9 atoms are stored as plain floats (nothing to decode); option "xtc_host_decode" keeps the
        // decoding reader threads for comparison
        src.slot_bytes = 12 * (size_t)x->info.natoms; src.layout = 1;
        src.fill = [x](int64_t frame, unsigned char *dst, double cell[9], FeedFill &, std::string &err) {
            thread_local std::vector<unsigned char> buf;
            return xtc_read_into(x, frame, buf, (float *)dst, cell, nullptr, nullptr, err);
        };
    } else {
        int64_t longest = 0;
        for (int64_t l : x->length) longest = std::max(longest, l);
        // worst case: every atom its own group (12 B of record each) + the longest compressed block of the file
        src.slot_bytes = sizeof(XtcDevHeader) + sizeof(XtcGroup) * (size_t)x->info.natoms + (size_t)longest + 64;
        src.decoded_bytes = 12 * (size_t)x->info.natoms; src.layout = 2;
        const size_t slot_bytes = src.slot_bytes;
        src.fill = [x, slot_bytes](int64_t frame, unsigned char *dst, double cell[9], FeedFill &out, std::string &err) {
            thread_local std::vector<unsigned char> buf;
            if (!xtc_read_into(x, frame, buf, nullptr, cell, nullptr, nullptr, err)) return false;
            int ng = 0;
            const size_t used = xtc_skeleton(buf.data() + XTC_HEADER, buf.size() - XTC_HEADER, (int)x->info.natoms, dst, slot_bytes, &ng);
            if (!used) { err = "malformed XTC coordinate block in frame " + std::to_string((long long)frame); return false; }
            out.ranges.push_back({0, used}); out.aux = ng;
            return true;
        };
    }
    // host decoding costs ~1.3 ms per 100 k atoms and thread, the skeleton walk of the device decoder a fraction of
    // that: by default half of the cores this process may run on, 2..8
    const int hw = (int)std::thread::hardware_concurrency();
    const int dflt = std::max(2, std::min(8, hw / 2));
    return run_feed(h, src, solute_indices, solvent_indices, frames, weights, nframes, n_reader_threads > 0 ? n_reader_threads : dflt);
}

}  // extern "C"
