// io.cpp — see io.hpp.  Whole-file reads + hand-rolled token scanners (std::from_chars, body lines parsed on all cores): the
// 871 k-triangle ASCII PLY of the dragon-scale config (~37 MB of text) parses in well under 0.1 s.
#include "io.hpp"

#include <algorithm>
#include <cctype>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace rtb {

namespace {
bool slurp(const char* path, std::string& out, std::string& err) {
    FILE* f = std::fopen(path, "rb");
    if (!f) { err = std::string("Couldn't open the file: ") + path; return false; }
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    const size_t got = n > 0 ? std::fread(&out[0], 1, (size_t)n, f) : 0;
    std::fclose(f);
    if (got != out.size()) { err = "short read"; return false; }
    return true;
}
// formats one channel like Rust's `{}` for f64: integers without ".0"
int fmt_channel(double v, char* b) {
    if (v == std::floor(v) && std::fabs(v) < 1e15) {
        long long i = (long long)v;
        if (i >= 0 && i < 1000) { // fast path for 0..255
            int n = 0;
            if (i >= 100) b[n++] = (char)('0' + i / 100);
            if (i >= 10) b[n++] = (char)('0' + (i / 10) % 10);
            b[n++] = (char)('0' + i % 10);
            return n;
        }
        return std::sprintf(b, "%lld", i);
    }
    return std::sprintf(b, "%.17g", v);
}
} // namespace

bool write_ppm_p3(const char* path, const double* screen, int32_t W, int32_t H, std::string& err) {
    std::string out;
    out.reserve((size_t)W * H * 12 + 32);
    char buf[96];
    out.append(buf, (size_t)std::sprintf(buf, "P3\n%d %d\n255\n", W, H));
    for (int32_t j = H - 1; j >= 0; --j) {
        const double* row = screen + (size_t)j * W * 3;
        for (int32_t i = 0; i < W; ++i) {
            int n = fmt_channel(row[3 * i], buf);
            buf[n++] = ' ';
            n += fmt_channel(row[3 * i + 1], buf + n);
            buf[n++] = ' ';
            n += fmt_channel(row[3 * i + 2], buf + n);
            buf[n++] = '\n';
            out.append(buf, (size_t)n);
        }
    }
    FILE* f = path ? std::fopen(path, "wb") : stdout;
    if (!f) { err = std::string("cannot open for writing: ") + path; return false; }
    const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
    if (path) std::fclose(f); else std::fflush(f);
    if (!ok) err = "short write";
    return ok;
}

// One decimal number at p (leading blanks and an optional '+' skipped, as Rust's str::parse::<f64> accepts them after
// split_whitespace, model.rs:44-48): std::from_chars is correctly rounded like strtod and several times faster; anything it does
// not take (hex floats, "infinity" spellings) goes through strtod, so the accepted language is the old one.
inline bool parse_double(const char*& p, const char* end, double& v) {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) ++p;
    const char* q = (p < end && *p == '+') ? p + 1 : p;
    const std::from_chars_result r = std::from_chars(q, end, v);
    if (r.ec == std::errc() && r.ptr != q) { p = r.ptr; return true; }
    // strtod knows no end pointer and skips '\n': run it on a bounded copy of the token, so that a line with too few numbers
    // is an error (the reference panics on such a file, model.rs:44-48) instead of borrowing a value from the next line
    const char* t = p;
    while (t < end && *t != ' ' && *t != '\t' && *t != '\r' && *t != '\n') ++t;
    if (t == p || t - p > 63) return false;
    char tok[64];
    std::memcpy(tok, p, (size_t)(t - p));
    tok[t - p] = 0;
    char* e;
    v = std::strtod(tok, &e);
    if (e == tok || *e != 0) return false;
    p = t;
    return true;
}
inline bool parse_ulong(const char*& p, const char* end, unsigned long& v) {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) ++p;
    const char* q = (p < end && *p == '+') ? p + 1 : p;
    const std::from_chars_result r = std::from_chars(q, end, v);
    if (r.ec != std::errc() || r.ptr == q) return false;
    p = r.ptr;
    return true;
}

bool read_ppm_p3(const char* path, int32_t& W, int32_t& H, std::vector<double>& rgb, std::string& err) {
    std::string s;
    if (!slurp(path, s, err)) return false;
    // line 0 = magic (ignored), line 1 = "W H" split on a single space, line 2 = maxval (ignored)
    const size_t l1 = s.find('\n');
    if (l1 == std::string::npos) { err = "ppm: missing size line"; return false; }
    const size_t l2 = s.find('\n', l1 + 1);
    if (l2 == std::string::npos) { err = "ppm: missing maxval line"; return false; }
    size_t l3 = s.find('\n', l2 + 1);
    if (l3 == std::string::npos) l3 = s.size();
    {
        const std::string wh = s.substr(l1 + 1, l2 - l1 - 1);
        const size_t sp = wh.find(' ');
        if (sp == std::string::npos) { err = "ppm: bad size line"; return false; }
        W = (int32_t)std::strtol(wh.c_str(), nullptr, 10);
        H = (int32_t)std::strtol(wh.c_str() + sp + 1, nullptr, 10);
    }
    if (W <= 0 || H <= 0) { err = "ppm: bad size"; return false; }
    rgb.clear();
    rgb.reserve((size_t)W * H * 3);
    const char* p = s.c_str() + std::min(l3 + 1, s.size());
    const char* end = s.c_str() + s.size();
    const size_t want = (size_t)W * H * 3;
    while (rgb.size() < want) {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r' || *p == '\f' || *p == '\v')) ++p;
        if (p >= end) break;
        double v;
        if (!parse_double(p, end, v)) { err = "ppm: bad sample"; return false; }
        rgb.push_back(v);
    }
    if (rgb.size() < want) { err = "ppm: not enough samples"; return false; }
    return true;
}

// ASCII PLY subset of TriangleModel::load_from_file (model.rs:13-62): "element vertex N" / "element face M" / "end_header",
// then N lines "x y z ..." (scaled) and M lines "3 a b c".  Lines are located with memchr, then parsed on all cores (each line is
// independent); errors are reported for the first bad line, as the serial reader did.
bool parse_ply_ascii(const std::string& s, double scale, std::vector<double>& verts, std::vector<uint32_t>& faces, std::string& err) {
    const char* p = s.c_str();
    const char* end = p + s.size();
    long nv = 0, nf = 0;
    bool header_done = false;
    while (p < end) {
        const char* nl = (const char*)std::memchr(p, '\n', (size_t)(end - p));
        const char* le = nl ? nl : end;
        if (le > p && le[-1] == '\r') --le; // CRLF files: read_ply_any's format probe tolerates the '\r', so does the header scan
        const size_t len = (size_t)(le - p);
        if (len == 10 && std::memcmp(p, "end_header", 10) == 0) { header_done = true; p = nl ? nl + 1 : end; break; }
        if (len > 15 && std::memcmp(p, "element vertex ", 15) == 0) nv = std::strtol(p + 15, nullptr, 10);
        if (len > 13 && std::memcmp(p, "element face ", 13) == 0) nf = std::strtol(p + 13, nullptr, 10);
        p = nl ? nl + 1 : end;
    }
    if (!header_done) { err = "ply: no end_header line"; return false; }
    if (nv < 0 || nf < 0) { err = "ply: negative counts"; return false; }
    // line starts of the nv + nf body lines
    const size_t n_lines = (size_t)nv + (size_t)nf;
    std::vector<const char*> line(n_lines + 1, end);
    size_t found = 0;
    while (found < n_lines && p < end) {
        line[found++] = p;
        const char* nl = (const char*)std::memchr(p, '\n', (size_t)(end - p));
        p = nl ? nl + 1 : end;
    }
    if (found < (size_t)nv) { err = "ply: truncated vertex list"; return false; }
    if (found < n_lines) { err = "ply: truncated face list"; return false; }
    line[n_lines] = p;
    verts.assign((size_t)nv * 3, 0.0);
    faces.assign((size_t)nf * 3, 0u);
    const unsigned hw = std::thread::hardware_concurrency();
    const unsigned threads = n_lines >= 65536 ? std::min(16u, std::max(1u, hw)) : 1u;
    // first failing line per kind (0 = bad vertex line, 1 = bad face line, 2 = index out of range); SIZE_MAX = none
    std::vector<size_t> bad(3 * (size_t)threads, SIZE_MAX);
    auto work = [&](unsigned t) {
        // every thread takes an equal share of the vertex lines AND of the face lines (vertex lines cost ~3x a face line)
        const size_t v_lo = (size_t)nv * t / threads, v_hi = (size_t)nv * (t + 1) / threads;
        const size_t f_lo = (size_t)nv + (size_t)nf * t / threads, f_hi = (size_t)nv + (size_t)nf * (t + 1) / threads;
        for (size_t i = v_lo; i < f_hi; i = (i + 1 == v_hi ? f_lo : i + 1)) {
            if (i >= v_hi && i < f_lo) { i = f_lo; if (i >= f_hi) break; }
            const char* q = line[i];
            const char* le = line[i + 1];
            if (i < (size_t)nv) {
                for (int k = 0; k < 3; ++k) {
                    double v;
                    if (!parse_double(q, le, v)) { bad[3 * t] = std::min(bad[3 * t], i); break; }
                    verts[3 * i + k] = v * scale;
                }
            } else {
                const size_t fi = i - (size_t)nv;
                unsigned long v;
                if (!parse_ulong(q, le, v)) { bad[3 * t + 1] = std::min(bad[3 * t + 1], i); continue; } // the leading count is ignored (model.rs:53-57)
                for (int k = 0; k < 3; ++k) {
                    if (!parse_ulong(q, le, v)) { bad[3 * t + 1] = std::min(bad[3 * t + 1], i); break; }
                    if (v >= (unsigned long)nv) { bad[3 * t + 2] = std::min(bad[3 * t + 2], i); break; }
                    faces[3 * fi + k] = (uint32_t)v;
                }
            }
        }
    };
    if (threads <= 1) work(0);
    else {
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < threads; ++t) pool.emplace_back(work, t);
        work(0);
        for (std::thread& th : pool) th.join();
    }
    size_t first = SIZE_MAX;
    int kind = -1;
    for (unsigned t = 0; t < threads; ++t)
        for (int k = 0; k < 3; ++k)
            if (bad[3 * t + k] < first) { first = bad[3 * t + k]; kind = k; }
    if (kind == 0) { err = "ply: bad vertex line"; return false; }
    if (kind == 1) { err = "ply: bad face line"; return false; }
    if (kind == 2) { err = "ply: vertex index out of range"; return false; }
    return true;
}

bool read_ply_ascii(const char* path, double scale, std::vector<double>& verts, std::vector<uint32_t>& faces, std::string& err) {
    std::string s;
    if (!slurp(path, s, err)) return false;
    return parse_ply_ascii(s, scale, verts, faces, err);
}

// ------------------------------------------------------------------ binary fast paths
bool write_ppm_p6(const char* path, const double* screen, int32_t W, int32_t H, std::string& err) {
    std::string out;
    char hdr[64];
    const int hn = std::sprintf(hdr, "P6\n%d %d\n255\n", W, H);
    out.resize((size_t)hn + (size_t)W * H * 3);
    std::memcpy(&out[0], hdr, (size_t)hn);
    unsigned char* q = reinterpret_cast<unsigned char*>(&out[0]) + hn;
    for (int32_t j = H - 1; j >= 0; --j) {
        const double* row = screen + (size_t)j * W * 3;
        for (int32_t i = 0; i < 3 * W; ++i) {
            const double v = row[i];
            *q++ = (unsigned char)(v <= 0.0 ? 0 : (v >= 255.0 ? 255 : (int)v));
        }
    }
    FILE* f = path ? std::fopen(path, "wb") : stdout;
    if (!f) { err = std::string("cannot open for writing: ") + path; return false; }
    const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
    if (path) std::fclose(f); else std::fflush(f);
    if (!ok) err = "short write";
    return ok;
}

bool read_ppm_any(const char* path, int32_t& W, int32_t& H, std::vector<double>& rgb, std::string& err) {
    {
        FILE* f = std::fopen(path, "rb");
        if (!f) { err = std::string("Couldn't open the file: ") + path; return false; }
        char magic[2] = {0, 0};
        const size_t got = std::fread(magic, 1, 2, f);
        std::fclose(f);
        if (got != 2 || magic[0] != 'P' || magic[1] != '6') return read_ppm_p3(path, W, H, rgb, err);
    }
    std::string s;
    if (!slurp(path, s, err)) return false;
    // header tokens: P6 W H maxval, separated by whitespace, '#' comments to end of line; one whitespace byte before the samples
    size_t p = 2;
    long vals[3] = {0, 0, 0};
    for (int k = 0; k < 3; ++k) {
        for (;;) {
            while (p < s.size() && std::isspace((unsigned char)s[p])) ++p;
            if (p < s.size() && s[p] == '#') { while (p < s.size() && s[p] != '\n') ++p; continue; }
            break;
        }
        char* e;
        vals[k] = std::strtol(s.c_str() + p, &e, 10);
        if (e == s.c_str() + p) { err = "ppm: bad P6 header"; return false; }
        p = (size_t)(e - s.c_str());
    }
    ++p;
    W = (int32_t)vals[0]; H = (int32_t)vals[1];
    if (W <= 0 || H <= 0 || vals[2] <= 0 || vals[2] > 255) { err = "ppm: unsupported P6 header"; return false; }
    const size_t want = (size_t)W * H * 3;
    if (s.size() < p + want) { err = "ppm: not enough samples"; return false; }
    rgb.resize(want);
    const unsigned char* q = reinterpret_cast<const unsigned char*>(s.data()) + p;
    for (size_t i = 0; i < want; ++i) rgb[i] = (double)q[i];
    return true;
}

namespace {
int ply_type_size(const std::string& t) {
    if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
    if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
    if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
    if (t == "double" || t == "float64") return 8;
    return 0;
}
bool ply_is_float(const std::string& t) { return t == "float" || t == "float32" || t == "double" || t == "float64"; }
std::vector<std::string> split_ws(const std::string& line) {
    std::vector<std::string> out;
    size_t i = 0;
    while (i < line.size()) {
        while (i < line.size() && std::isspace((unsigned char)line[i])) ++i;
        size_t j = i;
        while (j < line.size() && !std::isspace((unsigned char)line[j])) ++j;
        if (j > i) out.push_back(line.substr(i, j - i));
        i = j;
    }
    return out;
}
uint64_t load_uint(const unsigned char* p, int size) {
    uint64_t v = 0;
    for (int k = 0; k < size; ++k) v |= (uint64_t)p[k] << (8 * k); // little endian
    return v;
}
} // namespace

bool read_ply_any(const char* path, double scale, std::vector<double>& verts, std::vector<uint32_t>& faces, std::string& err) {
    std::string s;
    if (!slurp(path, s, err)) return false;
    // header
    size_t p = 0;
    bool binary = false, header_done = false;
    long nv = 0, nf = 0;
    int elem = 0; // 1 = vertex, 2 = face, 3 = other
    std::vector<std::pair<std::string, int>> vprops; // (type, size)
    int face_count_size = 0, face_index_size = 0, face_props = 0;
    while (p < s.size()) {
        size_t nl = s.find('\n', p);
        if (nl == std::string::npos) nl = s.size();
        std::string line = s.substr(p, nl - p);
        if (!line.empty() && line.back() == '\r') line.pop_back();
        p = nl + 1;
        const std::vector<std::string> tk = split_ws(line);
        if (tk.empty()) continue;
        if (tk[0] == "end_header") { header_done = true; break; }
        if (tk[0] == "format" && tk.size() >= 2) {
            if (tk[1] == "binary_little_endian") binary = true;
            else if (tk[1] != "ascii") { err = "ply: unsupported format " + tk[1]; return false; }
        } else if (tk[0] == "element" && tk.size() >= 3) {
            elem = tk[1] == "vertex" ? 1 : (tk[1] == "face" ? 2 : 3);
            if (elem == 1) nv = std::strtol(tk[2].c_str(), nullptr, 10);
            if (elem == 2) nf = std::strtol(tk[2].c_str(), nullptr, 10);
        } else if (tk[0] == "property") {
            if (elem == 1 && tk.size() >= 3) vprops.emplace_back(tk[1], ply_type_size(tk[1]));
            if (elem == 2) {
                ++face_props;
                if (tk.size() >= 5 && tk[1] == "list") { face_count_size = ply_type_size(tk[2]); face_index_size = ply_type_size(tk[3]); }
            }
        }
    }
    if (!header_done) { err = "ply: no end_header line"; return false; }
    if (!binary) return parse_ply_ascii(s, scale, verts, faces, err); // the buffer is already in memory
    if (nv < 0 || nf < 0) { err = "ply: negative counts"; return false; }
    if (vprops.size() < 3 || !ply_is_float(vprops[0].first) || !ply_is_float(vprops[1].first) || !ply_is_float(vprops[2].first)) {
        err = "ply: binary vertices need three leading float/double properties"; return false;
    }
    size_t vstride = 0;
    for (const auto& pr : vprops) { if (pr.second == 0) { err = "ply: unknown vertex property type " + pr.first; return false; } vstride += (size_t)pr.second; }
    if (face_props != 1 || face_count_size == 0 || face_index_size == 0) { err = "ply: binary faces need exactly one list property"; return false; }
    const unsigned char* q = reinterpret_cast<const unsigned char*>(s.data()) + p;
    const unsigned char* end = reinterpret_cast<const unsigned char*>(s.data()) + s.size();
    if ((size_t)(end - q) < vstride * (size_t)nv) { err = "ply: truncated vertex list"; return false; }
    verts.resize((size_t)nv * 3);
    for (long i = 0; i < nv; ++i) {
        const unsigned char* r = q + vstride * (size_t)i;
        for (int k = 0; k < 3; ++k) {
            double v;
            if (vprops[k].second == 4) { float f; std::memcpy(&f, r, 4); v = (double)f; } else { std::memcpy(&v, r, 8); }
            verts[3 * (size_t)i + k] = v * scale;
            r += vprops[k].second;
        }
    }
    q += vstride * (size_t)nv;
    faces.resize((size_t)nf * 3);
    for (long i = 0; i < nf; ++i) {
        if (q + face_count_size > end) { err = "ply: truncated face list"; return false; }
        const uint64_t cnt = load_uint(q, face_count_size);
        q += face_count_size;
        if (cnt != 3) { err = "ply: binary faces must be triangles"; return false; }
        if (q + 3 * (size_t)face_index_size > end) { err = "ply: truncated face list"; return false; }
        for (int k = 0; k < 3; ++k) {
            const uint64_t v = load_uint(q, face_index_size);
            q += face_index_size;
            if (v >= (uint64_t)nv) { err = "ply: vertex index out of range"; return false; }
            faces[3 * (size_t)i + k] = (uint32_t)v;
        }
    }
    return true;
}

bool write_ply_binary(const char* path, const std::vector<double>& verts, const std::vector<uint32_t>& faces, std::string& err) {
    FILE* f = std::fopen(path, "wb");
    if (!f) { err = std::string("cannot open for writing: ") + path; return false; }
    std::fprintf(f, "ply\nformat binary_little_endian 1.0\nelement vertex %zu\nproperty float x\nproperty float y\nproperty float z\nelement face %zu\n"
                    "property list uchar int vertex_indices\nend_header\n", verts.size() / 3, faces.size() / 3);
    std::vector<unsigned char> buf;
    buf.resize(verts.size() * 4 + faces.size() / 3 * 13);
    unsigned char* q = buf.data();
    for (double v : verts) { const float x = (float)v; std::memcpy(q, &x, 4); q += 4; }
    for (size_t i = 0; i + 2 < faces.size(); i += 3) {
        *q++ = 3;
        for (int k = 0; k < 3; ++k) { const int32_t v = (int32_t)faces[i + k]; std::memcpy(q, &v, 4); q += 4; }
    }
    const bool ok = std::fwrite(buf.data(), 1, buf.size(), f) == buf.size();
    std::fclose(f);
    if (!ok) err = "short write";
    return ok;
}

} // namespace rtb
