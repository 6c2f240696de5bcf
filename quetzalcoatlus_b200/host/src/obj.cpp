// Multi-threaded OBJ tokenizer with the value semantics of the reference's regex loader (obj/obj.cpp:9-175), restated
// as hand-written matchers:
//   * a line counts only if its type -- the run of non-blank characters from COLUMN 0 (`^\S+`, obj.cpp:131-139) -- is
//     exactly "v", "vn" or "f"; indented lines, comments and every other type are skipped;
//   * `v` / `vn`: the keyword, then blank-separated tokens (`\S+`), three required, a fourth optional for `v`; each token
//     goes through strtof as std::stof does (a numeric prefix is enough) (obj.cpp:9-40);
//   * `f`: the reference tries three UNANCHORED patterns in turn (obj.cpp:42-114) -- plain `a b c [d]`, `a/t ...`,
//     `a/[t]/n ...` -- so a pattern may match a prefix of the line and stop (`f 1 2 3/4` is the triangle 1 2 3, `f 1 2 3 4 5`
//     the quad 1 2 3 4), the fourth corner is taken only if it has the pattern's form, and a triangle repeats its third
//     corner.  The `a/[t]/n` form carries textures only if the first three corners all have one; where the reference
//     leaves `textures` unset (indeterminate values) this loader stores zeros.
// Where the reference throws (std::stof / std::stoi on a token without a number, or out of range) the record is dropped.
#include "../obj/obj.hpp"

#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <thread>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace obj {

namespace {

// `\s` inside a line ('\n' ends the line and never occurs in one)
inline bool is_blank(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }
inline bool is_end(char c) { return c == '\0' || c == '\n'; }
inline bool is_digit(char c) { return c >= '0' && c <= '9'; }
inline const char* skip_blank(const char* p) {
    while (is_blank(*p)) p++;
    return p;
}
inline const char* token_end(const char* p) {
    while (!is_end(*p) && !is_blank(*p)) p++;
    return p;
}

// strtof for the tokens OBJ writers produce -- [+-]digits[.digits], nothing else in the token -- when the result is
// exact by construction: the digits form an integer m < 2^24 and there are k <= 10 fraction digits, so m and 10^k are
// both floats and the one IEEE division m / 10^k rounds the decimal value once, as strtof does (Clinger's fast path).
// Anything else (exponents, more digits, inf/nan, hex, junk inside the token) returns false and goes to strtof.
bool exact_decimal(const char* p, float& value) {
    static const float pow10[11] = {1e0f, 1e1f, 1e2f, 1e3f, 1e4f, 1e5f, 1e6f, 1e7f, 1e8f, 1e9f, 1e10f};
    const bool neg = *p == '-';
    if (*p == '-' || *p == '+') p++;
    uint32_t m = 0;
    int digits = 0, frac = 0;
    for (; is_digit(*p); p++, digits++) {
        if (m > 1677720u) return false;
        m = m * 10u + uint32_t(*p - '0');
    }
    if (*p == '.') {
        for (p++; is_digit(*p); p++, digits++, frac++) {
            if (m > 1677720u) return false;
            m = m * 10u + uint32_t(*p - '0');
        }
    }
    if (digits == 0 || frac > 10 || m >= (1u << 24)) return false;
    if (!is_end(*p) && !is_blank(*p)) return false;
    const float v = float(m) / pow10[frac];
    value = neg ? -v : v;
    return true;
}

// std::stof on the token that starts at p: false where it would throw (no conversion)
bool token_float(const char* p, float& value) {
    if (exact_decimal(p, value)) return true;
    char* end = nullptr;
    value = std::strtof(p, &end);
    return end != p;
}

// `KEY\s+(\S+)\s+(\S+)\s+(\S+)(?:\s+(\S+))?` matched at p (behind the keyword): the tokens' starts; returns how many
int split_tokens(const char* p, const char* tok[], int max) {
    int n = 0;
    while (n < max && is_blank(*p)) {
        p = skip_blank(p);
        if (is_end(*p)) break;
        tok[n++] = p;
        p = token_end(p);
    }
    return n;
}

std::optional<Vertex> vertex_at(const char* p) {
    const char* tok[4];
    const int n = split_tokens(p, tok, 4);
    Vertex v{};
    if (n < 3 || !token_float(tok[0], v.x) || !token_float(tok[1], v.y) || !token_float(tok[2], v.z)) return std::nullopt;
    v.w = 1.0f;
    if (n == 4 && !token_float(tok[3], v.w)) return std::nullopt;
    return v;
}

std::optional<VertexNormal> normal_at(const char* p) {
    const char* tok[3];
    VertexNormal n{};
    if (split_tokens(p, tok, 3) < 3 || !token_float(tok[0], n.x) || !token_float(tok[1], n.y) || !token_float(tok[2], n.z)) return std::nullopt;
    return n;
}

// `(\d+)` as std::stoi reads it; values beyond int (where stoi throws) saturate and fail the caller's range check
inline bool digits(const char*& p, int& value) {
    if (!is_digit(*p)) return false;
    long long v = 0;
    for (; is_digit(*p); p++) v = std::min<long long>(v * 10 + (*p - '0'), INT_MAX);
    value = int(v);
    return true;
}

struct Corner {
    int v = 0, t = 0, n = 0;
    bool has_t = false;
};

// one corner of face pattern FORM at p: 0 `(\d+)`, 1 `(\d+)\/(\d+)`, 2 `(\d+)\/(\d+)?\/(\d+)`
template <int FORM>
inline bool corner(const char*& p, Corner& c) {
    if (!digits(p, c.v)) return false;
    if (FORM == 0) return true;
    if (*p++ != '/') return false;
    if (FORM == 1) return digits(p, c.t);
    c.has_t = digits(p, c.t);
    if (*p++ != '/') return false;
    return digits(p, c.n);
}

// `f\s+C\s+C\s+C(?:\s+C)?` with C the corner of FORM, matched at p (behind the 'f'): the number of corners or 0
template <int FORM>
int corners_at(const char* p, Corner c[4]) {
    for (int k = 0; k < 3; k++) {
        if (!is_blank(*p)) return 0;
        p = skip_blank(p);
        if (!corner<FORM>(p, c[k])) return 0;
    }
    if (!is_blank(*p)) return 3;
    p = skip_blank(p);
    return corner<FORM>(p, c[3]) ? 4 : 3;
}

// regex_search of pattern FORM over the line: the leftmost 'f' at which it matches
template <int FORM>
int search_corners(const char* line, Corner c[4]) {
    for (const char* p = line; !is_end(*p); p++) {
        if (*p != 'f') continue;
        const int n = corners_at<FORM>(p + 1, c);
        if (n) return n;
    }
    return 0;
}

std::optional<FaceElement> face_in(const char* line) {
    Corner c[4];
    FaceElement fe{};
    int n;
    if ((n = search_corners<0>(line, c))) {
        if (n == 3) c[3] = c[2];
        fe.vertices = {c[0].v, c[1].v, c[2].v, c[3].v};
    } else if ((n = search_corners<1>(line, c))) {
        if (n == 3) c[3] = c[2];
        fe.vertices = {c[0].v, c[1].v, c[2].v, c[3].v};
        fe.textures = {c[0].t, c[1].t, c[2].t, c[3].t};
    } else if ((n = search_corners<2>(line, c))) {
        const bool textured = c[0].has_t && c[1].has_t && c[2].has_t;
        const int t3 = (n == 4 && c[3].has_t) ? c[3].t : 0;   // (a triangle's fourth texture index is 0, not the third's)
        if (n == 3) c[3] = c[2];
        fe.vertices = {c[0].v, c[1].v, c[2].v, c[3].v};
        fe.normals = {c[0].n, c[1].n, c[2].n, c[3].n};
        if (textured) fe.textures = {c[0].t, c[1].t, c[2].t, t3};
    } else {
        return std::nullopt;
    }
    fe.n_vertices = size_t(n);
    return fe;
}

// regex_search of a `KEY\s+...` record pattern: the leftmost occurrence of the keyword at which the record parses
template <class R, class F>
std::optional<R> search_record(const char* line, const char* key, F at) {
    const size_t len = std::strlen(key);
    for (const char* p = line; !is_end(*p); p++) {
        if (std::strncmp(p, key, len) != 0) continue;
        if (auto r = at(p + len)) return r;
    }
    return std::nullopt;
}

// What a range of lines holds / where its records go: the counting pass only classifies the lines (an upper bound of
// the records -- a malformed line yields none), the parsing pass writes the records to v / vn / f and counts them.
struct Range {
    size_t nv = 0, nn = 0, nf = 0;
    Vertex* v = nullptr;
    VertexNormal* vn = nullptr;
    FaceElement* f = nullptr;
    bool has_nul = false;   // (pass 1) the range holds a NUL byte: see parse_text
};

// Lines of text[begin, end), `begin` at a line start.  The text ends in NUL at text[size] (the last line may lack its
// '\n').
template <bool PARSE>
void scan_lines(const char* text, size_t size, size_t begin, size_t end, Range& r) {
    size_t pos = begin, nv = 0, nn = 0, nf = 0;
    if (!PARSE) r.has_nul = end > begin && std::memchr(text + begin, '\0', end - begin) != nullptr;
    while (pos < end) {
        const char* p = text + pos;
        const void* nl = std::memchr(p, '\n', size - pos);
        pos = nl ? size_t(static_cast<const char*>(nl) - text) + 1 : size;
        if (is_end(*p) || is_blank(*p) || *p == '#') continue;
        const char* e = token_end(p);
        const size_t len = size_t(e - p);
        if (len == 1 && p[0] == 'v') {
            if (!PARSE) nv++;
            else if (auto v = vertex_at(e)) r.v[nv++] = *v;
        } else if (len == 2 && p[0] == 'v' && p[1] == 'n') {
            if (!PARSE) nn++;
            else if (auto vn = normal_at(e)) r.vn[nn++] = *vn;
        } else if (len == 1 && p[0] == 'f') {
            if (!PARSE) nf++;
            else if (auto fc = face_in(p)) r.f[nf++] = *fc;
        }
    }
    r.nv = nv, r.nn = nn, r.nf = nf;
}

std::string without_nul(const char* text, size_t size) {
    std::string copy(text, size);
    std::replace(copy.begin(), copy.end(), '\0', '\x01');
    return copy;
}

template <class F>
void run_parts(size_t n_parts, F f) {
    std::vector<std::thread> workers;
    for (size_t k = 1; k < n_parts; k++) workers.emplace_back([&f, k] { f(k); });
    f(0);
    for (auto& w : workers) w.join();
}

// The text is cut at line boundaries into one range per host thread; a record depends on nothing but its own line
// (indices are absolute), so the ranges are independent.  Pass 1 counts each range's lines by type, which places the
// range's records in the final arrays; pass 2 parses straight into them.  Records of malformed lines leave gaps at the
// end of a range's share, closed afterwards in file order.
ObjData parse_text(const char* text, size_t size) {
    const size_t n_parts = std::max<size_t>(1, std::min<size_t>({std::thread::hardware_concurrency(), 16, size / (1u << 20) + 1}));
    std::vector<size_t> cut(n_parts + 1, size);
    cut[0] = 0;
    for (size_t k = 1; k < n_parts; k++) {
        const size_t from = std::max(cut[k - 1], size / n_parts * k);
        const void* nl = from < size ? std::memchr(text + from, '\n', size - from) : nullptr;
        cut[k] = nl ? size_t(static_cast<const char*>(nl) - text) + 1 : size;
    }
    std::vector<Range> count(n_parts), done(n_parts);
    run_parts(n_parts, [&](size_t k) { scan_lines<false>(text, size, cut[k], cut[k + 1], count[k]); });
    // A NUL byte INSIDE the text is an ordinary non-blank character to the reference (`\S` matches it; std::stof stops at
    // it like at any other non-numeric character), but this parser's end-of-text sentinel.  Such a file is parsed from a
    // copy in which every NUL is the byte 0x01, which the reference's patterns and strtof treat the same way.
    for (const Range& c : count)
        if (c.has_nul) return parse_text(without_nul(text, size).c_str(), size);
    size_t nv = 0, nn = 0, nf = 0;
    for (const Range& c : count) nv += c.nv, nn += c.nn, nf += c.nf;
    ObjData data;
    data.vertices.resize(nv);
    data.vertex_normals.resize(nn);
    data.faces.resize(nf);
    nv = nn = nf = 0;
    for (size_t k = 0; k < n_parts; k++) {
        done[k].v = data.vertices.data() + nv, done[k].vn = data.vertex_normals.data() + nn, done[k].f = data.faces.data() + nf;
        nv += count[k].nv, nn += count[k].nn, nf += count[k].nf;
    }
    run_parts(n_parts, [&](size_t k) { scan_lines<true>(text, size, cut[k], cut[k + 1], done[k]); });
    auto close_gaps = [&](auto& vec, auto ptr, auto num) {
        size_t out = 0;
        for (size_t k = 0; k < n_parts; k++) {
            const size_t at = size_t(done[k].*ptr - vec.data()), n = done[k].*num;
            if (at != out) std::move(vec.begin() + at, vec.begin() + at + n, vec.begin() + out);
            out += n;
        }
        vec.resize(out);
    };
    close_gaps(data.vertices, &Range::v, &Range::nv);
    close_gaps(data.vertex_normals, &Range::vn, &Range::nn);
    close_gaps(data.faces, &Range::f, &Range::nf);
    return data;
}

}  // namespace

// (the reference's per-record entry points, obj.hpp:12-33: regex_search, so the record may start anywhere in the line)
// (a std::string may hold NUL bytes: same substitution as in parse_text)
std::optional<Vertex> Vertex::from_line(const std::string& line) {
    if (line.find('\0') != std::string::npos) return from_line(without_nul(line.data(), line.size()));
    return search_record<Vertex>(line.c_str(), "v", vertex_at);
}

std::optional<VertexNormal> VertexNormal::from_line(const std::string& line) {
    if (line.find('\0') != std::string::npos) return from_line(without_nul(line.data(), line.size()));
    return search_record<VertexNormal>(line.c_str(), "vn", normal_at);
}

std::optional<FaceElement> FaceElement::from_line(const std::string& line) {
    if (line.find('\0') != std::string::npos) return from_line(without_nul(line.data(), line.size()));
    return face_in(line.c_str());
}

// The file is mapped, not copied: the parser reads the page cache in place.  It needs a NUL behind the last byte, which
// the zero fill of the mapping's last page provides -- except for a file that fills its last page exactly and does not
// end in a newline, which is read into a string instead.
std::optional<ObjData> load_obj(const std::string& filename) {
    std::cerr << "Loading " << filename << "..." << std::endl;
    const int fd = ::open(filename.c_str(), O_RDONLY);
    struct stat st{};
    if (fd < 0 || ::fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) {
        if (fd >= 0) ::close(fd);
        std::cout << "Unable to open file " << filename << std::endl;
        return std::nullopt;
    }
    const size_t size = size_t(st.st_size);
    if (size == 0) {
        ::close(fd);
        return ObjData{};
    }
    void* map = ::mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    const size_t page = size_t(::sysconf(_SC_PAGESIZE));
    std::optional<ObjData> data;
    if (map != MAP_FAILED && (size % page != 0 || static_cast<const char*>(map)[size - 1] == '\n')) {
        ::madvise(map, size, MADV_WILLNEED);
        data = parse_text(static_cast<const char*>(map), size);
    } else {
        std::string text(size, '\0');
        size_t got = 0;
        for (ssize_t n; got < size && (n = ::pread(fd, text.data() + got, size - got, off_t(got))) > 0;) got += size_t(n);
        text.resize(got);
        data = parse_text(text.c_str(), text.size());
    }
    if (map != MAP_FAILED) ::munmap(map, size);
    ::close(fd);
    return data;
}

}  // namespace obj
