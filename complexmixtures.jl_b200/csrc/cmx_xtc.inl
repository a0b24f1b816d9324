// cmx_xtc.inl -- native GROMACS XTC reader (host code, included by cmx_b200.cu): SURVEY 8 (f1), second format.
//
// The reference reads XTC through Chemfiles (src/trajectory_formats/ChemFiles.jl:112-138: read_step, positions in
// Angstrom, unit cell matrix); this is a from-the-format-description decoder of the XTC frame layout and of its
// compressed coordinate block (the "xdr3dfcoord" scheme: coordinates quantised to integers at `precision` per nm, the
// first atom of a run stored in full range with a mixed-radix packing of the three components, the following atoms
// as small differences in an adaptive range; the first two atoms of a run are swapped, which favours water).
// Output: fp32 xyz triplets in Angstrom (nm x 10 in double, rounded once to fp32) and the cell as the column-major 3x3
// matrix cmx_submit_frame takes.  Frames are indexed at open time so that any frame can be read by number and ranks can
// read disjoint frames.  Decoding on the device is the next step; today the decoder feeds the pinned staging slot.

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

}  // namespace

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

// The frame loop for an XTC file: like cmx_run_dcd, but the reader threads also DECODE (the compressed block is a
// serial bit stream per frame, so frames are the unit of parallelism: one frame per thread at a time); the ring
// holds decoded fp32 xyz triplets of the whole frame, the selection gather runs on the device.
int32_t cmx_run_xtc(cmx_handle *h, cmx_xtc *x, const int32_t *solute_indices, const int32_t *solvent_indices,
                    const int64_t *frames, const double *weights, int64_t nframes, int32_t n_reader_threads) {
    if (!h) return CMX_ERR_ARG;
    if (!x) return fail(h, CMX_ERR_ARG, "cmx_run_xtc: null argument");
    FeedSource src;
    src.what = "cmx_run_xtc"; src.natoms = x->info.natoms; src.nframes = x->info.nframes;
    src.slot_bytes = 12 * (size_t)x->info.natoms; src.layout = 1;
    src.fill = [x](int64_t frame, unsigned char *dst, double cell[9], std::string &err) {
        thread_local std::vector<unsigned char> buf;
        return xtc_read_into(x, frame, buf, (float *)dst, cell, nullptr, nullptr, err);
    };
    // decoding costs ~1.3 ms per 100 k atoms and thread: by default half of the cores this process may run on, 2..8
    const int hw = (int)std::thread::hardware_concurrency();
    const int dflt = std::max(2, std::min(8, hw / 2));
    return run_feed(h, src, solute_indices, solvent_indices, frames, weights, nframes, n_reader_threads > 0 ? n_reader_threads : dflt);
}

}  // extern "C"
