// Mesh construction helpers: Wavefront OBJ reader, axis-aligned planes and boxes.  Host-side scene set-up.
//
// Behavioural contract = the reference's src/scene/mesh.cpp, re-implemented over an
// in-memory byte buffer with flat arrays instead of a character stream and per-vertex vectors:
//   * records start after leading blanks; only "v " and "f " records are read, everything else is skipped to the end
//     of its line; a record does not consume the rest of its line beyond its last token's delimiter;
//   * a number token is the maximal run of [0-9 + - e E] (plus '.' for reals); the byte that ends the run is consumed
//     with it; conversion failures give NaN (reals) or -1 (integers);
//   * faces are 1-based and use the first three indices; "a//n" is tolerated because the '/' that ends `a` is eaten
//     with the token and the following "/n" group is skipped;
//   * faces with an out-of-range index, two coincident vertices or a zero cross product are dropped.
#include <PathTrace/scene/mesh.h>

#include <cerrno>
#include <climits>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <limits>
#include <string>
#include <vector>

namespace {

    class ObjReader {
      public:
        ObjReader(const std::string &text, const mat4<float> &transform, bool cull, bool smooth) :
          data(text.data()), size(text.size()), transform(transform), cull(cull), smooth(smooth) {}

        std::vector<Triangle> run() {
            while(more()) {
                record();
            }
            if(smooth) {
                smoothNormals();
            }
            return std::move(faces);
        }

      private:
        const char *data;
        std::size_t size;
        std::size_t at = 0;
        mat4<float> transform;
        bool cull;
        bool smooth;

        std::vector<vec3<float>> vertices;
        std::vector<Triangle> faces;
        std::vector<uint32_t> corner_vertex; // 3 per kept face: which vertex each corner refers to

        bool more() const { return at < size; }
        char look() const { return more() ? data[at] : static_cast<char>(-1); }
        char take() { return more() ? data[at++] : static_cast<char>(-1); }

        bool accept(char c) {
            if(more() && data[at] == c) {
                at++;
                return true;
            }
            return false;
        }

        void blanks() {
            while(accept(' ')) {
            }
        }

        void restOfLine() {
            while(more()) {
                const char c = take();
                if(c == '\r' || c == '\n') {
                    return;
                }
            }
        }

        // token = maximal run of number characters; the terminating byte is consumed as well
        std::string token(bool real) {
            blanks();
            const std::size_t begin = at;
            std::size_t end = at;
            while(more()) {
                const char c = take();
                const bool number_char = (c >= '0' && c <= '9') || c == '-' || c == '+' || c == 'e' || c == 'E' || (real && c == '.');
                if(!number_char) {
                    break;
                }
                end = at;
            }
            return std::string(data + begin, end - begin);
        }

        float real() {
            const std::string word = token(true);
            char *stop = nullptr;
            errno = 0;
            const float value = std::strtof(word.c_str(), &stop);
            if(stop == word.c_str() || errno == ERANGE) {
                return std::numeric_limits<float>::quiet_NaN();
            }
            return value;
        }

        int integer() {
            const std::string word = token(false);
            char *stop = nullptr;
            errno = 0;
            const long value = std::strtol(word.c_str(), &stop, 10);
            if(stop == word.c_str() || errno == ERANGE || value < INT_MIN || value > INT_MAX) {
                return -1;
            }
            return static_cast<int>(value);
        }

        int faceIndex() {
            const int index = integer() - 1;
            while(accept('/')) {
                integer();
            }
            return index;
        }

        void vertexRecord() {
            const float x = real();
            const float y = real();
            const float z = real();
            vertices.emplace_back(transform * vec3<float>{x, y, z});
        }

        void faceRecord() {
            const int ia = faceIndex();
            const int ib = faceIndex();
            const int ic = faceIndex();
            const int count = static_cast<int>(vertices.size());
            if(ia < 0 || ia >= count || ib < 0 || ib >= count || ic < 0 || ic >= count) {
                return;
            }
            const vec3<float> &a = vertices[ia];
            const vec3<float> &b = vertices[ib];
            const vec3<float> &c = vertices[ic];
            // written so that NaN coordinates fail the test
            const bool distinct = (b - a).getLengthSquared() > 0.0F && (c - a).getLengthSquared() > 0.0F && (c - b).getLengthSquared() > 0.0F;
            if(!distinct) {
                return;
            }
            if(cross(b - a, c - a).getLengthSquared() <= 0.0F) {
                return;
            }
            faces.emplace_back(a, b, c, cull);
            corner_vertex.push_back(static_cast<uint32_t>(ia));
            corner_vertex.push_back(static_cast<uint32_t>(ib));
            corner_vertex.push_back(static_cast<uint32_t>(ic));
        }

        void record() {
            blanks();
            switch(take()) {
                case '\r':
                case '\n':
                    break;
                case 'v':
                    if(accept(' ')) {
                        vertexRecord();
                    }
                    else {
                        restOfLine();
                    }
                    break;
                case 'f':
                    if(accept(' ')) {
                        faceRecord();
                    }
                    else {
                        restOfLine();
                    }
                    break;
                default: // comments and every other record type
                    restOfLine();
                    break;
            }
        }

        // vertex normal = normalised sum of the unit face normals of all incident faces, accumulated in face order
        void smoothNormals() {
            const std::size_t n_faces = faces.size();
            std::vector<vec3<float>> unit_normal(n_faces);
            for(std::size_t f = 0; f < n_faces; f++) {
                unit_normal[f] = cross(faces[f].b - faces[f].a, faces[f].c - faces[f].a).normalize();
            }

            // incidence lists in CSR form (counting sort keeps face order within a vertex)
            std::vector<uint32_t> first(vertices.size() + 1, 0U);
            for(uint32_t v : corner_vertex) {
                first[v + 1]++;
            }
            for(std::size_t v = 0; v < vertices.size(); v++) {
                first[v + 1] += first[v];
            }
            std::vector<uint32_t> corners(corner_vertex.size());
            {
                std::vector<uint32_t> cursor(first.begin(), first.end() - 1);
                for(uint32_t corner = 0; corner < corner_vertex.size(); corner++) {
                    corners[cursor[corner_vertex[corner]]++] = corner;
                }
            }

            for(std::size_t v = 0; v < vertices.size(); v++) {
                vec3<float> sum{};
                for(uint32_t k = first[v]; k < first[v + 1]; k++) {
                    sum = sum + unit_normal[corners[k] / 3U];
                }
                if(sum.getLengthSquared() <= 0.0F) {
                    continue;
                }
                const vec3<float> normal = sum.normalize();
                for(uint32_t k = first[v]; k < first[v + 1]; k++) {
                    Triangle &face = faces[corners[k] / 3U];
                    switch(corners[k] % 3U) {
                        case 0:
                            face.normal_a = normal;
                            break;
                        case 1:
                            face.normal_b = normal;
                            break;
                        default:
                            face.normal_c = normal;
                            break;
                    }
                }
            }
        }
    };

}

namespace io {

    std::vector<Triangle> loadMesh(std::basic_istream<char> &stream, mat4<float> transformation, bool cull_backface, bool smooth) {
        const std::string text((std::istreambuf_iterator<char>(stream)), std::istreambuf_iterator<char>());
        return ObjReader(text, transformation, cull_backface, smooth).run();
    }

    std::vector<Triangle> loadMesh(const std::filesystem::path &path, mat4<float> transformation, bool cull_backface, bool smooth) {
        std::ifstream stream(path, std::ios_base::in | std::ios_base::binary);
        if(!stream) {
            return {};
        }
        return loadMesh(stream, transformation, cull_backface, smooth);
    }

}

std::vector<Triangle> makePlane(vec3<float> a, vec3<float> b, bool cull_backface) {
    constexpr float tolerance = 1E-4F;

    // the plane's normal axis is the LAST axis along which the corners coincide; exactly one axis may coincide
    int flat_axis = -1;
    int flat_count = 0;
    for(int axis = 0; axis < 3; axis++) {
        if(std::abs(a[axis] - b[axis]) < tolerance) {
            flat_axis = axis;
            flat_count++;
        }
    }
    if(flat_count != 1) {
        return {};
    }

    // the two remaining corners swap the coordinate of the first in-plane axis
    const int swap_axis = flat_axis == 0 ? 1 : 0;
    vec3<float> corner_ab = a;
    vec3<float> corner_ba = b;
    corner_ab[swap_axis] = b[swap_axis];
    corner_ba[swap_axis] = a[swap_axis];

    std::vector<Triangle> triangles;
    triangles.reserve(2);
    triangles.emplace_back(a, corner_ab, b, cull_backface);
    triangles.emplace_back(b, corner_ba, a, cull_backface);
    return triangles;
}

std::vector<Triangle> makeBox(vec3<float> a, vec3<float> b, bool cull_backface) {
    constexpr float tolerance = 1E-4F;
    for(int axis = 0; axis < 3; axis++) {
        if(std::abs(a[axis] - b[axis]) < tolerance) {
            return {};
        }
    }

    std::vector<Triangle> triangles;
    triangles.reserve(12);
    for(int axis = 0; axis < 3; axis++) {
        // the face through a[axis], then the opposite face through b[axis]
        for(const float level : {a[axis], b[axis]}) {
            vec3<float> lo = a;
            vec3<float> hi = b;
            lo[axis] = level;
            hi[axis] = level;
            const auto face = makePlane(lo, hi, cull_backface);
            triangles.insert(triangles.end(), face.begin(), face.end());
        }
    }
    return triangles;
}
