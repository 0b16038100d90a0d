// irec_io.cpp -- host C++ of include/irec_io.h: the integer arithmetic coder and the `.rec` container of the
// reference's index stream (rec/io/entropy_coding.pyx, rec/io/utils.py).  No CUDA calls.
//
// Interval arithmetic follows the reference exactly (same comparisons, same strictness: `high < half or low > half`,
// `low > quarter and high < 3 * quarter`, final `low <= quarter`), so code strings are bit-identical.  Products
// width * mass are formed in 128 bits (the reference's encode uses Python integers; its decode C longs -- both exact
// for precision <= 32 and total mass < 2^31).
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <vector>

#include "../../include/irec_io.h"
#include "irec_host.h"

typedef unsigned __int128 u128;

namespace {

struct Model {
    std::vector<int64_t> C, Dc;   // cumulative masses before / after each symbol (entropy_coding.pyx:27-47)
    int64_t R;
    bool ok;
};

Model make_model(const int64_t* counts, int n)
{
    Model m;
    m.ok = n > 0;
    m.C.resize(n > 0 ? n : 0); m.Dc.resize(n > 0 ? n : 0);
    int64_t c = 0;
    for (int i = 0; i < n; ++i) {
        if (counts[i] <= 0) m.ok = false;      // zero-mass symbols make the reference's interval tree loop forever
        m.C[i] = c; c += counts[i]; m.Dc[i] = c;
    }
    m.R = c;
    if (c <= 0 || c >= (1LL << 62)) m.ok = false;
    return m;
}

inline int64_t scale(int64_t width, int64_t mass, int64_t R) { return (int64_t)(((u128)width * (u128)mass) / (u128)R); }

struct BitSink {
    uint8_t* out; int64_t cap; int64_t n;
    void put(int bit) { if (n < cap) out[n] = (uint8_t)bit; ++n; }
    void put_run(int first, int s) { put(first); for (int i = 0; i < s; ++i) put(1 - first); }
};

int ac_encode(const Model& m, int precision, const int64_t* msg, int64_t n_msg, BitSink& sink)
{
    const int64_t whole = 1LL << precision, half = whole >> 1, quarter = whole >> 2;
    int64_t low = 0, high = whole;
    int s = 0;
    const int64_t ns = (int64_t)m.C.size();
    for (int64_t k = 0; k < n_msg; ++k) {
        const int64_t sym = msg[k];
        if (sym < 0 || sym >= ns) return irec_fail(IREC_E_INVALID, "ac_encode: symbol outside the alphabet (index > max_index?)");
        const int64_t width = high - low;
        high = low + scale(width, m.Dc[sym], m.R);
        low = low + scale(width, m.C[sym], m.R);
        while (high < half || low > half) {                       // entropy_coding.pyx:84-101
            if (high < half) { sink.put_run(0, s); s = 0; low *= 2; high *= 2; }
            else { sink.put_run(1, s); s = 0; low = (low - half) * 2; high = (high - half) * 2; }
        }
        while (low > quarter && high < 3 * quarter) {             // :104-107
            ++s; low = (low - quarter) * 2; high = (high - quarter) * 2;
        }
        if (high <= low) return irec_fail(IREC_E_INVALID, "ac_encode: interval collapsed (precision too small for these masses)");
    }
    ++s;                                                          // :110-115
    if (low <= quarter) sink.put_run(0, s); else sink.put_run(1, s);
    return IREC_OK;
}

int ac_decode(const Model& m, int precision, const uint8_t* bits, int64_t n_bits, std::vector<int64_t>& out)
{
    const int64_t whole = 1LL << precision, half = whole >> 1, quarter = whole >> 2;
    int64_t low = 0, high = whole, z = 0, i = 0;
    while (i < precision && i < n_bits) {                         // :232-235
        if (bits[i]) z += 1LL << (precision - i - 1);
        ++i;
    }
    const int ns = (int)m.C.size();
    // every symbol costs at least one interval update; a valid code of n bits cannot hold more than this many symbols
    // unless masses are extreme -- the bound only stops runaway decoding of corrupt input
    const int64_t max_symbols = 64 * (n_bits + 64) + 1024;
    for (;;) {
        const int64_t width = high - low, target = z - low;
        if (target < 0 || width <= 0) return irec_fail(IREC_E_INVALID, "ac_decode: code does not decode (lower bound below the interval)");
        // largest j with floor(width * C[j] / R) <= target   (data_structures.py:184-210)
        int lo = 0, hi = ns - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (scale(width, m.C[mid], m.R) <= target) lo = mid; else hi = mid - 1;
        }
        const int j = lo;
        const int64_t low_ = low + scale(width, m.C[j], m.R), high_ = low + scale(width, m.Dc[j], m.R);
        if (!(z < high_)) return irec_fail(IREC_E_INVALID, "ac_decode: code does not decode (no symbol interval contains it)");
        out.push_back(j);
        high = high_; low = low_;
        if (j == 0) return IREC_OK;                               // :262-266
        if ((int64_t)out.size() > max_symbols) return irec_fail(IREC_E_INVALID, "ac_decode: no end-of-message symbol");
        while (high < half || low > half) {                       // :269-286
            if (high < half) { low *= 2; high *= 2; z *= 2; }
            else { low = (low - half) * 2; high = (high - half) * 2; z = (z - half) * 2; }
            if (i < n_bits && bits[i]) z += 1;
            ++i;
        }
        while (low > quarter && high < 3 * quarter) {             // :289-298
            low = (low - quarter) * 2; high = (high - quarter) * 2; z = (z - quarter) * 2;
            if (i < n_bits && bits[i]) z += 1;
            ++i;
        }
    }
}

// '1' + code as a big-endian integer in ceil(len/8) bytes (rec/io/utils.py:66-72,100-106)
void bits_to_bytes(const std::vector<uint8_t>& code, std::vector<uint8_t>& out)
{
    const int64_t L = (int64_t)code.size() + 1, nbytes = (L + 7) / 8, pad = nbytes * 8 - L;
    const size_t base = out.size();
    out.resize(base + (size_t)nbytes, 0);
    for (int64_t b = 0; b < L; ++b) {
        const int bit = b == 0 ? 1 : code[(size_t)(b - 1)];
        if (bit) out[base + (size_t)((pad + b) >> 3)] |= (uint8_t)(0x80u >> ((pad + b) & 7));
    }
}

// inverse: bin(int.from_bytes(...))[3:]  (rec/io/utils.py:151-164)
bool bytes_to_bits(const uint8_t* p, int64_t nbytes, std::vector<uint8_t>& code)
{
    code.clear();
    bool seen = false;
    for (int64_t b = 0; b < nbytes * 8; ++b) {
        const int bit = (p[b >> 3] >> (7 - (b & 7))) & 1;
        if (!seen) { seen = bit != 0; continue; }
        code.push_back((uint8_t)bit);
    }
    return seen;
}

std::vector<int64_t> default_index_counts(uint32_t max_index)
{
    std::vector<int64_t> c((size_t)max_index + 1, 1001);          // rec/io/utils.py:31-35
    c[0] = 1;
    return c;
}
std::vector<int64_t> default_nav_counts(int64_t nav_max)
{
    std::vector<int64_t> c((size_t)nav_max + 2, 101);             // rec/io/utils.py:43-49
    c[0] = 1;
    return c;
}

int encode_stream(const std::vector<int64_t>& counts, const int64_t* vals, int64_t n, std::vector<uint8_t>& code)
{
    // to_message: values + 1, then the end symbol 0  (rec/io/utils.py:58-59)
    std::vector<int64_t> msg((size_t)n + 1);
    for (int64_t i = 0; i < n; ++i) msg[(size_t)i] = vals[i] + 1;
    msg[(size_t)n] = 0;
    const Model m = make_model(counts.data(), (int)counts.size());
    if (!m.ok) return irec_fail(IREC_E_INVALID, "rec_pack: bad symbol masses");
    code.assign((size_t)(n + 2) * 40 + 128, 0);
    for (;;) {
        BitSink sink{ code.data(), (int64_t)code.size(), 0 };
        const int rc = ac_encode(m, 32, msg.data(), (int64_t)msg.size(), sink);
        if (rc != IREC_OK) return rc;
        if (sink.n <= (int64_t)code.size()) { code.resize((size_t)sink.n); return IREC_OK; }
        code.assign((size_t)sink.n, 0);
    }
}

const int64_t STATIC_HEADER = 28;     // struct.calcsize('IIIIIHHHH')

void put_u32(std::vector<uint8_t>& o, uint32_t v) { for (int i = 0; i < 4; ++i) o.push_back((uint8_t)(v >> (8 * i))); }
void put_u16(std::vector<uint8_t>& o, uint16_t v) { for (int i = 0; i < 2; ++i) o.push_back((uint8_t)(v >> (8 * i))); }
uint32_t get_u32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint16_t get_u16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

int pack_image(const irec_rec_header_t* h, const int32_t* num_blocks, const int32_t* num_aux, const int64_t* indices,
               const int64_t* index_counts, std::vector<uint8_t>& file)
{
    if (!h || h->n_res_blocks < 0 || (h->n_res_blocks > 0 && (!num_blocks || !num_aux)))
        return irec_fail(IREC_E_INVALID, "rec_pack: bad arguments");
    const int nr = h->n_res_blocks;
    std::vector<int64_t> icounts = index_counts ? std::vector<int64_t>(index_counts, index_counts + h->max_index + 1)
                                                : default_index_counts(h->max_index);
    std::vector<std::vector<uint8_t>> nav_bytes((size_t)nr), idx_bytes((size_t)nr);
    std::vector<uint32_t> nav_max((size_t)nr);
    int64_t blk0 = 0, idx0 = 0;
    std::vector<uint8_t> code;
    for (int r = 0; r < nr; ++r) {
        const int nbk = num_blocks[r];
        if (nbk <= 0) return irec_fail(IREC_E_INVALID, "rec_pack: a latent tensor without coder-blocks (the reference's np.max fails on it)");
        std::vector<int64_t> nav((size_t)nbk);
        int64_t mx = 0, n_idx = 0;
        for (int b = 0; b < nbk; ++b) {
            nav[(size_t)b] = num_aux[blk0 + b];
            if (nav[(size_t)b] < 0) return irec_fail(IREC_E_INVALID, "rec_pack: negative index count");
            mx = std::max(mx, nav[(size_t)b]);
            n_idx += nav[(size_t)b];
        }
        nav_max[(size_t)r] = (uint32_t)mx;
        int rc = encode_stream(default_nav_counts(mx), nav.data(), nbk, code);
        if (rc != IREC_OK) return rc;
        bits_to_bytes(code, nav_bytes[(size_t)r]);
        rc = encode_stream(icounts, indices + idx0, n_idx, code);
        if (rc != IREC_OK) return rc;
        bits_to_bytes(code, idx_bytes[(size_t)r]);
        blk0 += nbk; idx0 += n_idx;
    }
    file.clear();
    put_u32(file, h->seed); put_u32(file, h->block_size); put_u32(file, h->max_index);
    put_u32(file, h->image_h); put_u32(file, h->image_w);
    put_u16(file, h->image_c); put_u16(file, 0); put_u16(file, index_counts ? 1 : 0); put_u16(file, (uint16_t)nr);
    for (int r = 0; r < nr; ++r) put_u32(file, (uint32_t)num_blocks[r]);
    for (int r = 0; r < nr; ++r) put_u32(file, (uint32_t)nav_bytes[(size_t)r].size());
    for (int r = 0; r < nr; ++r) put_u32(file, (uint32_t)idx_bytes[(size_t)r].size());
    for (int r = 0; r < nr; ++r) put_u32(file, nav_max[(size_t)r]);
    for (int r = 0; r < nr; ++r) file.insert(file.end(), nav_bytes[(size_t)r].begin(), nav_bytes[(size_t)r].end());
    for (int r = 0; r < nr; ++r) file.insert(file.end(), idx_bytes[(size_t)r].begin(), idx_bytes[(size_t)r].end());
    return IREC_OK;
}

}  // namespace

extern "C" {

int irec_ac_encode(const int64_t* counts, int n_symbols, int precision, const int64_t* message, int64_t n_message,
                   uint8_t* out_bits, int64_t capacity, int64_t* out_n_bits)
{
    if (!counts || n_symbols <= 0 || (!message && n_message > 0) || n_message < 0 || !out_n_bits || capacity < 0 ||
        (!out_bits && capacity > 0))
        return irec_fail(IREC_E_INVALID, "ac_encode: bad arguments");
    if (precision < 3 || precision > 60) return irec_fail(IREC_E_INVALID, "ac_encode: precision must be in [3, 60]");
    const Model m = make_model(counts, n_symbols);
    if (!m.ok) return irec_fail(IREC_E_INVALID, "ac_encode: symbol masses must be positive");
    BitSink sink{ out_bits, capacity, 0 };
    const int rc = ac_encode(m, precision, message, n_message, sink);
    if (rc != IREC_OK) return rc;
    *out_n_bits = sink.n;
    return sink.n <= capacity ? IREC_OK : irec_fail(IREC_E_CAPACITY, "ac_encode: output buffer too small");
}

int irec_ac_decode(const int64_t* counts, int n_symbols, int precision, const uint8_t* bits, int64_t n_bits,
                   int64_t* out_message, int64_t capacity, int64_t* out_n_message)
{
    if (!counts || n_symbols <= 0 || (!bits && n_bits > 0) || n_bits < 0 || !out_n_message || capacity < 0 ||
        (!out_message && capacity > 0))
        return irec_fail(IREC_E_INVALID, "ac_decode: bad arguments");
    if (precision < 3 || precision > 60) return irec_fail(IREC_E_INVALID, "ac_decode: precision must be in [3, 60]");
    const Model m = make_model(counts, n_symbols);
    if (!m.ok) return irec_fail(IREC_E_INVALID, "ac_decode: symbol masses must be positive");
    std::vector<int64_t> out;
    const int rc = ac_decode(m, precision, bits, n_bits, out);
    if (rc != IREC_OK) return rc;
    *out_n_message = (int64_t)out.size();
    const int64_t n = std::min<int64_t>((int64_t)out.size(), capacity);
    if (n > 0) memcpy(out_message, out.data(), sizeof(int64_t) * (size_t)n);
    return (int64_t)out.size() <= capacity ? IREC_OK : irec_fail(IREC_E_CAPACITY, "ac_decode: output buffer too small");
}

int irec_rec_pack(const irec_rec_header_t* header, const int32_t* num_blocks, const int32_t* num_aux, const int64_t* indices,
                  const int64_t* index_counts, uint8_t* out, int64_t capacity, int64_t* out_bytes)
{
    if (!out_bytes || capacity < 0 || (!out && capacity > 0)) return irec_fail(IREC_E_INVALID, "rec_pack: bad arguments");
    std::vector<uint8_t> file;
    const int rc = pack_image(header, num_blocks, num_aux, indices, index_counts, file);
    if (rc != IREC_OK) return rc;
    *out_bytes = (int64_t)file.size();
    if ((int64_t)file.size() > capacity) return irec_fail(IREC_E_CAPACITY, "rec_pack: output buffer too small");
    if (!file.empty()) memcpy(out, file.data(), file.size());
    return IREC_OK;
}

int irec_rec_write_file(const char* path, const irec_rec_header_t* header, const int32_t* num_blocks, const int32_t* num_aux,
                        const int64_t* indices, const int64_t* index_counts, int64_t* out_bytes)
{
    if (!path) return irec_fail(IREC_E_INVALID, "rec_write_file: no path");
    std::vector<uint8_t> file;
    const int rc = pack_image(header, num_blocks, num_aux, indices, index_counts, file);
    if (rc != IREC_OK) return rc;
    FILE* f = fopen(path, "wb");
    if (!f) return irec_fail(IREC_E_INVALID, "rec_write_file: cannot open the file for writing");
    const size_t w = file.empty() ? 0 : fwrite(file.data(), 1, file.size(), f);
    const int closed = fclose(f);
    if (w != file.size() || closed != 0) return irec_fail(IREC_E_INVALID, "rec_write_file: short write");
    if (out_bytes) *out_bytes = (int64_t)file.size();
    return IREC_OK;
}

int irec_rec_read_header(const uint8_t* file, int64_t file_bytes, irec_rec_header_t* h)
{
    if (!file || !h) return irec_fail(IREC_E_INVALID, "rec_read_header: bad arguments");
    if (file_bytes < STATIC_HEADER) return irec_fail(IREC_E_INVALID, "rec_read_header: file shorter than the static header");
    h->seed = get_u32(file); h->block_size = get_u32(file + 4); h->max_index = get_u32(file + 8);
    h->image_h = get_u32(file + 12); h->image_w = get_u32(file + 16); h->image_c = get_u16(file + 20);
    h->uses_num_aux_counts_file = get_u16(file + 22); h->uses_index_counts_file = get_u16(file + 24);
    h->n_res_blocks = (int32_t)get_u16(file + 26);
    return IREC_OK;
}

int irec_rec_unpack(const uint8_t* file, int64_t file_bytes, const int64_t* index_counts, int32_t* num_blocks,
                    int32_t* num_aux, int64_t num_aux_capacity, int64_t* out_n_num_aux,
                    int64_t* indices, int64_t indices_capacity, int64_t* out_n_indices)
{
    irec_rec_header_t h;
    int rc = irec_rec_read_header(file, file_bytes, &h);
    if (rc != IREC_OK) return rc;
    if (!num_blocks || !out_n_num_aux || !out_n_indices) return irec_fail(IREC_E_INVALID, "rec_unpack: bad arguments");
    if (h.uses_index_counts_file && !index_counts)
        return irec_fail(IREC_E_INVALID, "rec_unpack: the file uses empirical index counts, but none were supplied");   // utils.py:137-138
    if (h.uses_num_aux_counts_file)
        return irec_fail(IREC_E_INVALID, "rec_unpack: empirical num_aux_var counts are not supported (the reference cannot write such files: struct.pack('I', -1))");
    const int nr = h.n_res_blocks;
    const int64_t dyn = 16LL * nr;
    if (file_bytes < STATIC_HEADER + dyn) return irec_fail(IREC_E_INVALID, "rec_unpack: truncated dynamic header");
    const uint8_t* d = file + STATIC_HEADER;
    std::vector<uint32_t> nav_len((size_t)nr), idx_len((size_t)nr), nav_max((size_t)nr);
    int64_t payload = 0;
    for (int r = 0; r < nr; ++r) {
        num_blocks[r] = (int32_t)get_u32(d + 4 * r);
        nav_len[(size_t)r] = get_u32(d + 4 * (nr + r));
        idx_len[(size_t)r] = get_u32(d + 4 * (2 * nr + r));
        nav_max[(size_t)r] = get_u32(d + 4 * (3 * nr + r));
        payload += (int64_t)nav_len[(size_t)r] + idx_len[(size_t)r];
    }
    if (file_bytes < STATIC_HEADER + dyn + payload) return irec_fail(IREC_E_INVALID, "rec_unpack: truncated code section");
    const std::vector<int64_t> icounts = index_counts ? std::vector<int64_t>(index_counts, index_counts + h.max_index + 1)
                                                      : default_index_counts(h.max_index);
    const Model im = make_model(icounts.data(), (int)icounts.size());
    if (!im.ok) return irec_fail(IREC_E_INVALID, "rec_unpack: bad index masses");
    const uint8_t* p_nav = file + STATIC_HEADER + dyn;
    const uint8_t* p_idx = p_nav;
    for (int r = 0; r < nr; ++r) p_idx += nav_len[(size_t)r];
    int64_t n_nav = 0, n_idx = 0;
    std::vector<uint8_t> code;
    std::vector<int64_t> msg;
    for (int r = 0; r < nr; ++r) {
        if (!bytes_to_bits(p_nav, nav_len[(size_t)r], code)) return irec_fail(IREC_E_INVALID, "rec_unpack: empty code");
        const std::vector<int64_t> nc = default_nav_counts(nav_max[(size_t)r]);
        const Model nm = make_model(nc.data(), (int)nc.size());
        msg.clear();
        rc = ac_decode(nm, 32, code.data(), (int64_t)code.size(), msg);
        if (rc != IREC_OK) return rc;
        int64_t expect = 0;
        for (size_t i = 0; i + 1 < msg.size(); ++i) {             // from_message: drop the end symbol, subtract 1
            if (n_nav < num_aux_capacity) num_aux[n_nav] = (int32_t)(msg[i] - 1);
            ++n_nav;
            expect += msg[i] - 1;
        }
        if ((int64_t)msg.size() - 1 != num_blocks[r]) return irec_fail(IREC_E_INVALID, "rec_unpack: block count does not match the header");
        p_nav += nav_len[(size_t)r];
        if (!bytes_to_bits(p_idx, idx_len[(size_t)r], code)) return irec_fail(IREC_E_INVALID, "rec_unpack: empty code");
        msg.clear();
        rc = ac_decode(im, 32, code.data(), (int64_t)code.size(), msg);
        if (rc != IREC_OK) return rc;
        if ((int64_t)msg.size() - 1 != expect) return irec_fail(IREC_E_INVALID, "rec_unpack: index count does not match the auxiliary-variable counts");
        for (size_t i = 0; i + 1 < msg.size(); ++i) {
            if (n_idx < indices_capacity) indices[n_idx] = msg[i] - 1;
            ++n_idx;
        }
        p_idx += idx_len[(size_t)r];
    }
    *out_n_num_aux = n_nav; *out_n_indices = n_idx;
    if (n_nav > num_aux_capacity || n_idx > indices_capacity) return irec_fail(IREC_E_CAPACITY, "rec_unpack: output buffers too small");
    return IREC_OK;
}

}  // extern "C"
