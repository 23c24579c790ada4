// Triangle-mesh construction: OBJ loading and the plane / box helpers
// (API of the reference's include/PathTrace/scene/mesh.h).  Host-side scene set-up, not accelerated.
#ifndef PATHTRACE_MESH_H
#define PATHTRACE_MESH_H

#include <PathTrace/scene/object.h>

#include <filesystem>
#include <istream>
#include <memory>
#include <vector>

namespace io {

    /**
     * Reads `v x y z` and `f a b c` records (1-based positive indices, `a//n` tolerated) from Wavefront OBJ text.
     * Vertices are transformed (with w-divide) as they are read; faces with out-of-range indices, coincident
     * vertices or zero area are dropped; with `smooth`, vertex normals are the normalised sum of the unit face
     * normals of all incident faces.
     */
    std::vector<Triangle> loadMesh(std::basic_istream<char> &stream, mat4<float> transformation = mat4_identity<float>, bool cull_backface = true,
                                   bool smooth = true);

    //! as above, reading the file at `path`; a missing file yields an empty vector
    std::vector<Triangle> loadMesh(const std::filesystem::path &path, mat4<float> transformation = mat4_identity<float>, bool cull_backface = true,
                                   bool smooth = true);

}

//! Axis-aligned rectangle with diagonal (a, b) as two triangles; empty for arguments that do not span a rectangle
std::vector<Triangle> makePlane(vec3<float> a, vec3<float> b, bool cull_backface = false);

//! Axis-aligned box with space diagonal (a, b) as twelve triangles; empty if the box is flat in any dimension
std::vector<Triangle> makeBox(vec3<float> a, vec3<float> b, bool cull_backface = false);

//! Copies plain objects into owning pointers appended to `objects`
template<typename T>
void moveObjects(std::vector<std::unique_ptr<Object>> &objects, std::vector<T> &extension) {
    objects.reserve(objects.size() + extension.size());
    for(auto &item : extension) {
        objects.emplace_back(std::make_unique<T>(item));
    }
}

#endif // PATHTRACE_MESH_H
