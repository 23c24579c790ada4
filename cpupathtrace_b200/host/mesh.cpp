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

#include <algorithm>
#include <array>
#include <cerrno>
#include <chrono>
#include <cstdio>
#include <climits>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <limits>
#include <string>
#include <thread>
#include <vector>

namespace {

    class ObjReader {
      public:
        ObjReader(const std::string &text, const mat4<float> &transform, bool cull, bool smooth) :
          data(text.data()), size(text.size()), transform(transform), cull(cull), smooth(smooth) {}

        static double nowSeconds() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
        static void phase(const char *what, double &since) {
            if(std::getenv("PTB_LOG_MESH") != nullptr) {
                const double now = nowSeconds();
                std::fprintf(stderr, "[mesh] %-28s %7.1f ms\n", what, (now - since) * 1e3);
                since = now;
            }
        }

        std::vector<Triangle> run() {
            double since = nowSeconds();
            unsigned workers = std::min(16U, std::max(1U, std::thread::hardware_concurrency()));
            if(const char *forced = std::getenv("PTB_MESH_THREADS")) {
                workers = static_cast<unsigned>(std::min(64L, std::max(1L, std::atol(forced))));
            }
            if(workers > 1U && size >= (1U << 20)) {
                runChunks(workers);
            }
            else {
                while(more()) {
                    record();
                }
            }
            phase("records -> faces", since);
            if(smooth) {
                smoothNormals();
            }
            phase("smooth normals", since);
            return std::move(faces);
        }

        // Large files are parsed in chunks on several threads.  The grammar above is sequential -- a record may end in the
        // middle of a line or spill into the next one (a vertex record with two numbers eats the first token of the
        // following line), and a face is kept only if its indices are below the number of vertices read SO FAR -- so a
        // chunk (cut after a line break) is parsed as if the stream started there, records where it stopped, and is only
        // accepted if the chunk before it stopped exactly at its first byte; from the first chunk that does not line up the
        // rest of the file is parsed sequentially.  Faces are validated afterwards against the vertex count at their position
        // in the stream, which the chunks' prefix sums give.  Result: identical to the sequential parse, byte for byte.
        struct PendingFace {
            int ia, ib, ic;
            uint32_t vertices_before; // within the chunk
        };

        void runChunks(unsigned workers) {
            std::vector<std::size_t> cut(workers + 1, size);
            cut[0] = 0;
            for(unsigned k = 1; k < workers; k++) {
                std::size_t p = std::max(cut[k - 1], size * k / workers);
                while(p < size && data[p] != '\n') {
                    p++;
                }
                cut[k] = std::min(size, p + 1);
            }
            std::vector<ObjReader> parts;
            parts.reserve(workers);
            for(unsigned k = 0; k < workers; k++) {
                parts.emplace_back(*this);
                parts.back().at = cut[k];
                parts.back().deferred = true;
            }
            {
                std::vector<std::thread> pool;
                for(unsigned k = 0; k < workers; k++) {
                    pool.emplace_back([&parts, &cut, k]() {
                        ObjReader &part = parts[k];
                        while(part.at < cut[k + 1] && part.more()) {
                            part.record();
                        }
                    });
                }
                for(std::thread &t : pool) {
                    t.join();
                }
            }
            double since = nowSeconds();
            phase("  (chunk parse done)", since);
            // accept the chunks that line up; parse the rest sequentially (in deferred mode as well) from where the last
            // accepted chunk stopped
            unsigned accepted = 1;
            while(accepted < workers && parts[accepted - 1].at == cut[accepted]) {
                accepted++;
            }
            if(parts[accepted - 1].at < size && accepted < workers) {
                ObjReader &tail = parts[accepted];
                tail.vertices.clear();
                tail.pending.clear();
                tail.at = parts[accepted - 1].at;
                while(tail.more()) {
                    tail.record();
                }
                accepted++;
            }
            parts.erase(parts.begin() + static_cast<std::ptrdiff_t>(accepted), parts.end());

            // vertices: concatenate; faces: validate against the vertex count at their position, in parallel per chunk
            std::vector<std::size_t> vertex_base(parts.size() + 1, 0);
            for(std::size_t k = 0; k < parts.size(); k++) {
                vertex_base[k + 1] = vertex_base[k] + parts[k].vertices.size();
            }
            vertices.resize(vertex_base.back());
            {
                std::vector<std::thread> pool;
                for(std::size_t k = 0; k < parts.size(); k++) {
                    pool.emplace_back([this, &parts, &vertex_base, k]() { std::copy(parts[k].vertices.begin(), parts[k].vertices.end(), vertices.begin() + static_cast<std::ptrdiff_t>(vertex_base[k])); });
                }
                for(std::thread &t : pool) {
                    t.join();
                }
            }
            {
                std::vector<std::thread> pool;
                for(std::size_t k = 0; k < parts.size(); k++) {
                    pool.emplace_back([this, &parts, &vertex_base, k]() {
                        ObjReader &part = parts[k];
                        part.faces.reserve(part.pending.size());
                        part.corner_vertex.reserve(3 * part.pending.size());
                        for(const PendingFace &f : part.pending) {
                            part.keepFace(vertices, static_cast<int>(vertex_base[k] + f.vertices_before), f.ia, f.ib, f.ic);
                        }
                    });
                }
                for(std::thread &t : pool) {
                    t.join();
                }
            }
            phase("  vertices + face validation", since);
            std::size_t total = 0;
            for(const ObjReader &part : parts) {
                total += part.faces.size();
            }
            faces.reserve(total);
            corner_vertex.reserve(3 * total);
            for(ObjReader &part : parts) {
                faces.insert(faces.end(), part.faces.begin(), part.faces.end());
                corner_vertex.insert(corner_vertex.end(), part.corner_vertex.begin(), part.corner_vertex.end());
            }
            at = size;
            phase("  concatenation", since);
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
        bool deferred = false;               // chunk mode: faces are recorded, not validated (runChunks)
        std::vector<PendingFace> pending;

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
            if(deferred) {
                pending.push_back(PendingFace{ia, ib, ic, static_cast<uint32_t>(vertices.size())});
                return;
            }
            keepFace(vertices, static_cast<int>(vertices.size()), ia, ib, ic);
        }

        // `count` = vertices read before the face record; `all` may hold more (chunk mode)
        void keepFace(const std::vector<vec3<float>> &all, int count, int ia, int ib, int ic) {
            if(ia < 0 || ia >= count || ib < 0 || ib >= count || ic < 0 || ic >= count) {
                return;
            }
            const vec3<float> &a = all[ia];
            const vec3<float> &b = all[ib];
            const vec3<float> &c = all[ic];
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
        const double t0 = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
        // the whole stream in blocks (an istreambuf_iterator copy moves one character per virtual call: 0.15 s for 40 MB)
        std::string text;
        {
            std::streambuf *buffer = stream.rdbuf();
            std::size_t filled = 0;
            for(;;) {
                if(text.size() - filled < (1U << 20)) {
                    text.resize(std::max<std::size_t>(2 * text.size(), 4U << 20));
                }
                const std::streamsize got = buffer != nullptr ? buffer->sgetn(&text[filled], static_cast<std::streamsize>(text.size() - filled)) : 0;
                if(got <= 0) {
                    break;
                }
                filled += static_cast<std::size_t>(got);
            }
            text.resize(filled);
            stream.setstate(std::ios_base::eofbit);
        }
        if(std::getenv("PTB_LOG_MESH") != nullptr) {
            std::fprintf(stderr, "[mesh] %-28s %7.1f ms\n", "stream -> memory", (std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count() - t0) * 1e3);
        }
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
