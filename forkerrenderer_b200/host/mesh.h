// mesh.h — Mesh of the drop-in facade (reference src/mesh.h:16-62, mesh.cpp:10-83).
// Holds the per-face index arrays; Draw(shader) is the draw-call site.  In the reference Draw loops over the
// faces on the CPU (vertex program x3 + ForkerGL::DrawTriangle per face); here it enqueues ONE indexed draw of
// the whole mesh on the device (fgl_draw_mesh), primitive ids following the same face order.
#pragma once

#include <memory>
#include <vector>

#include "geometry.h"
#include "material.h"

struct Shader;
class Model;

class Mesh : public std::enable_shared_from_this<Mesh>
{
public:
    explicit Mesh(const Model& model) : m_Model(model) {}
    Mesh(const Mesh&) = delete;

    void Draw(Shader& shader) const;

    int      NumFaces() const { return (int)m_FaceVertIndices.size() / 3; }
    Vector3f Vert(int faceIdx, int vertIdx) const;
    Vector2f TexCoord(int faceIdx, int vertIdx) const;
    Vector3f Normal(int faceIdx, int vertIdx) const;
    Vector3f Tangent(int faceIdx, int vertIdx) const;
    int      GetVertIndex(int faceIdx, int vertIdx) const { return m_FaceVertIndices[faceIdx * 3 + vertIdx]; }

    const Model&                       GetModel() const { return m_Model; }
    std::shared_ptr<const Material>    GetMaterial() const { return m_Material.lock(); }
    std::shared_ptr<const PBRMaterial> GetPBRMaterial() const { return m_PBRMaterial.lock(); }
    void SetMaterial(std::shared_ptr<const Material> m) { m_Material = m; }
    void SetPBRMaterial(std::shared_ptr<const PBRMaterial> m) { m_PBRMaterial = m; }

    void AddVertIndex(int i) { m_FaceVertIndices.push_back(i); }
    void AddTexCoordIndex(int i) { m_FaceTexCoordIndices.push_back(i); }
    void AddNormalIndex(int i) { m_FaceNormalIndices.push_back(i); }
    void AddTangentIndex(int i) { m_FaceTangentIndices.push_back(i); }

    const std::vector<int>& VertIndices() const { return m_FaceVertIndices; }
    const std::vector<int>& TexCoordIndices() const { return m_FaceTexCoordIndices; }
    const std::vector<int>& NormalIndices() const { return m_FaceNormalIndices; }

    int DeviceId() const;  // fgl mesh handle (uploads the owning model on first call)

private:
    friend class Model;
    const Model&                     m_Model;
    std::weak_ptr<const Material>    m_Material;
    std::weak_ptr<const PBRMaterial> m_PBRMaterial;
    std::vector<int>                 m_FaceVertIndices, m_FaceTexCoordIndices, m_FaceNormalIndices,
        m_FaceTangentIndices;
    mutable int m_DeviceId = -1;
};
