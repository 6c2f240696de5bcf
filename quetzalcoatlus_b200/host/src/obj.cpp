// Single-pass OBJ tokenizer with the value semantics of the reference's regex loader
// (obj/obj.cpp:9-175).  What the reference's three face patterns accept and how they fill
// FaceElement is kept: plain `f a b c [d]`; `f a/t b/t c/t [d/t]`; `f a/[t]/n ...`; at most
// four corners; a triangle repeats its third corner (obj.cpp:54-57,78-82,100-104); numbers
// are parsed with strtof/strtol (the reference uses std::stof/std::stoi on the same tokens).
#include "../obj/obj.hpp"

#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

namespace obj {

namespace {

const char* skip_ws(const char* p) {
    while (*p == ' ' || *p == '\t' || *p == '\r') p++;
    return p;
}

// parses up to `max` whitespace-separated floats after the keyword; returns how many
int parse_floats(const char* p, float* out, int max) {
    int n = 0;
    while (n < max) {
        p = skip_ws(p);
        if (!*p || *p == '\n') break;
        char* end = nullptr;
        float v = std::strtof(p, &end);
        if (end == p) break;
        out[n++] = v;
        p = end;
        // std::stof on a \S+ token ignores trailing junk in the token: skip to the next blank
        while (*p && *p != ' ' && *p != '\t' && *p != '\r' && *p != '\n') p++;
    }
    return n;
}

struct Corner {
    int v = 0, t = 0, n = 0;
    int slashes = 0;
    bool has_t = false, has_n = false;
};

// one face corner: digits [ '/' [digits] [ '/' digits ] ]
bool parse_corner(const char*& p, Corner& c) {
    p = skip_ws(p);
    if (*p < '0' || *p > '9') return false;
    char* end = nullptr;
    c.v = int(std::strtol(p, &end, 10));
    p = end;
    if (*p == '/') {
        c.slashes = 1;
        p++;
        if (*p >= '0' && *p <= '9') {
            c.t = int(std::strtol(p, &end, 10));
            c.has_t = true;
            p = end;
        }
        if (*p == '/') {
            c.slashes = 2;
            p++;
            if (*p >= '0' && *p <= '9') {
                c.n = int(std::strtol(p, &end, 10));
                c.has_n = true;
                p = end;
            }
        }
    }
    return true;
}

}  // namespace

std::optional<Vertex> Vertex::from_line(const std::string& line) {
    const char* p = std::strstr(line.c_str(), "v");
    if (!p) return std::nullopt;
    float f[4];
    int n = parse_floats(p + 1, f, 4);
    if (n < 3) return std::nullopt;
    Vertex v{f[0], f[1], f[2]};
    if (n == 4) v.w = f[3];
    return v;
}

std::optional<VertexNormal> VertexNormal::from_line(const std::string& line) {
    const char* p = std::strstr(line.c_str(), "vn");
    if (!p) return std::nullopt;
    float f[3];
    if (parse_floats(p + 2, f, 3) < 3) return std::nullopt;
    return VertexNormal{f[0], f[1], f[2]};
}

std::optional<FaceElement> FaceElement::from_line(const std::string& line) {
    const char* p = std::strstr(line.c_str(), "f");
    if (!p) return std::nullopt;
    p++;
    Corner c[4];
    int n = 0;
    while (n < 4) {
        const char* q = p;
        if (!parse_corner(q, c[n])) break;
        // corners must be blank-separated and of one form
        if (*q && *q != ' ' && *q != '\t' && *q != '\r' && *q != '\n') break;
        if (n > 0 && (c[n].slashes != c[0].slashes)) break;
        p = q;
        n++;
    }
    if (n < 3) return std::nullopt;
    FaceElement fe{};
    fe.n_vertices = size_t(n);
    const int form = c[0].slashes;
    if (form == 1) {
        for (int i = 0; i < n; i++) if (!c[i].has_t) return std::nullopt;  // `a/` is not a valid corner
    }
    if (form == 2) {
        for (int i = 0; i < n; i++) if (!c[i].has_n) return std::nullopt;
    }
    for (int i = 0; i < 4; i++) {
        const Corner& s = c[i < n ? i : n - 1];
        fe.vertices[i] = s.v;
        if (form == 1) fe.textures[i] = s.t;
        if (form == 2) fe.normals[i] = s.n;
    }
    if (form == 2 && c[0].has_t && c[1].has_t && c[2].has_t) {
        fe.textures = {c[0].t, c[1].t, c[2].t, (n == 4 && c[3].has_t) ? c[3].t : 0};
    }
    return fe;
}

std::optional<ObjData> load_obj(const std::string& filename) {
    std::cerr << "Loading " << filename << "..." << std::endl;
    std::ifstream file(filename, std::ios::binary);
    if (!file.is_open()) {
        std::cout << "Unable to open file " << filename << std::endl;
        return std::nullopt;
    }
    std::stringstream ss;
    ss << file.rdbuf();
    const std::string text = ss.str();
    ObjData data;
    size_t pos = 0;
    std::string line;
    while (pos < text.size()) {
        size_t eol = text.find('\n', pos);
        if (eol == std::string::npos) eol = text.size();
        line.assign(text, pos, eol - pos);
        pos = eol + 1;
        // the line type is the first run of non-blank characters; '#' starts a comment
        const char* p = line.c_str();
        while (*p == ' ' || *p == '\t' || *p == '\r') p++;
        if (!*p || *p == '#') continue;
        const char* e = p;
        while (*e && *e != ' ' && *e != '\t' && *e != '\r') e++;
        size_t len = size_t(e - p);
        if (len == 1 && p[0] == 'v') {
            if (auto v = Vertex::from_line(line)) data.vertices.push_back(*v);
        } else if (len == 2 && p[0] == 'v' && p[1] == 'n') {
            if (auto vn = VertexNormal::from_line(line)) data.vertex_normals.push_back(*vn);
        } else if (len == 1 && p[0] == 'f') {
            if (auto fc = FaceElement::from_line(line)) data.faces.push_back(*fc);
        }
    }
    return data;
}

}  // namespace obj
