// Wavefront OBJ loader with the reference's ObjData semantics (obj/obj.hpp:8-45,
// obj.cpp:9-175): 1-based indices kept as read, triangles stored as degenerate quads
// (4th index = 3rd), `vn` optional, face forms `a b c [d]`, `a/t ...`, `a/t/n ...`, `a//n ...`,
// at most four corners per face, unknown line types ignored.  The reference parses with one
// std::regex construction per line (minutes for a 1 M-face mesh); this is a single-pass
// strtof/strtol tokenizer producing the same values.
#pragma once

#include <array>
#include <optional>
#include <string>
#include <vector>

namespace obj {

struct Vertex {
    float x, y, z;
    float w = 1.0;
    static std::optional<Vertex> from_line(const std::string& line);
};

struct VertexNormal {
    float x, y, z;
    static std::optional<VertexNormal> from_line(const std::string& line);
};

struct FaceElement {
    std::array<int, 4> vertices;
    std::array<int, 4> textures;
    std::array<int, 4> normals;
    size_t n_vertices;
    static std::optional<FaceElement> from_line(const std::string& line);
};

struct ObjData {
    std::vector<Vertex> vertices;
    std::vector<VertexNormal> vertex_normals;
    std::vector<FaceElement> faces;
};

std::optional<ObjData> load_obj(const std::string& filename);

}  // namespace obj
