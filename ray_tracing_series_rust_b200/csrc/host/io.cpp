// io.cpp — see io.hpp.  Whole-file reads + hand-rolled token scanners: the 871 k-triangle ASCII PLY
// of the dragon-scale config (~40 MB of text) parses in well under a second.
#include "io.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

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
        char* e;
        const double v = std::strtod(p, &e);
        if (e == p) { err = "ppm: bad sample"; return false; }
        rgb.push_back(v);
        p = e;
    }
    if (rgb.size() < want) { err = "ppm: not enough samples"; return false; }
    return true;
}

bool read_ply_ascii(const char* path, double scale, std::vector<double>& verts, std::vector<uint32_t>& faces, std::string& err) {
    std::string s;
    if (!slurp(path, s, err)) return false;
    const char* p = s.c_str();
    const char* end = p + s.size();
    long nv = 0, nf = 0;
    bool header_done = false;
    while (p < end) {
        const char* nl = (const char*)std::memchr(p, '\n', (size_t)(end - p));
        const char* le = nl ? nl : end;
        const size_t len = (size_t)(le - p);
        if (len == 10 && std::memcmp(p, "end_header", 10) == 0) { header_done = true; p = nl ? nl + 1 : end; break; }
        if (len > 15 && std::memcmp(p, "element vertex ", 15) == 0) nv = std::strtol(p + 15, nullptr, 10);
        if (len > 13 && std::memcmp(p, "element face ", 13) == 0) nf = std::strtol(p + 13, nullptr, 10);
        p = nl ? nl + 1 : end;
    }
    if (!header_done) { err = "ply: no end_header line"; return false; }
    if (nv < 0 || nf < 0) { err = "ply: negative counts"; return false; }
    verts.clear();
    verts.reserve((size_t)nv * 3);
    for (long i = 0; i < nv; ++i) {
        if (p >= end) { err = "ply: truncated vertex list"; return false; }
        const char* nl = (const char*)std::memchr(p, '\n', (size_t)(end - p));
        char* e;
        for (int k = 0; k < 3; ++k) {
            const double v = std::strtod(p, &e);
            if (e == p) { err = "ply: bad vertex line"; return false; }
            verts.push_back(v * scale);
            p = e;
        }
        p = nl ? nl + 1 : end;
    }
    faces.clear();
    faces.reserve((size_t)nf * 3);
    for (long i = 0; i < nf; ++i) {
        if (p >= end) { err = "ply: truncated face list"; return false; }
        const char* nl = (const char*)std::memchr(p, '\n', (size_t)(end - p));
        char* e;
        (void)std::strtoul(p, &e, 10); // the leading count is ignored (model.rs:53-57)
        if (e == p) { err = "ply: bad face line"; return false; }
        p = e;
        for (int k = 0; k < 3; ++k) {
            const unsigned long v = std::strtoul(p, &e, 10);
            if (e == p) { err = "ply: bad face line"; return false; }
            if (v >= (unsigned long)nv) { err = "ply: vertex index out of range"; return false; }
            faces.push_back((uint32_t)v);
            p = e;
        }
        p = nl ? nl + 1 : end;
    }
    return true;
}

} // namespace rtb
