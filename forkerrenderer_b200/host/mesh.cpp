// mesh.cpp — Mesh accessors and the draw-call site (reference src/mesh.cpp:10-61).
#include "mesh.h"

#include "forkergl.h"
#include "model.h"
#include "shader.h"

// Reference Mesh::Draw (mesh.cpp:10-25): per face { shader.Use; 3x ProcessVertex; ForkerGL::DrawTriangle }.
// Default here: one indexed draw of all faces, vertex programs included, on the device.  With
// ForkerGL::SetPerTriangleSubmission(true) the reference's loop runs as it stands (host vertex programs, batched DrawTriangle).
void Mesh::Draw(Shader& shader) const
{
    if (!ForkerGL::GetPerTriangleSubmission())
    {
        shader.Use(shared_from_this());
        ForkerGL::DrawMesh(*this, shader);
        return;
    }
    for (int f = 0; f < NumFaces(); ++f)
    {
        shader.Use(shared_from_this());
        Point4f ndcCoords[3];
        for (int v = 0; v < 3; ++v) ndcCoords[v] = shader.ProcessVertex(f, v);
        ForkerGL::DrawTriangle(ndcCoords, shader);
    }
}

Vector3f Mesh::Vert(int faceIdx, int vertIdx) const
{
    return m_Model.GetVert(m_FaceVertIndices[faceIdx * 3 + vertIdx]);
}
Vector2f Mesh::TexCoord(int faceIdx, int vertIdx) const
{
    return m_Model.GetTexCoord(m_FaceTexCoordIndices[faceIdx * 3 + vertIdx]);
}
Vector3f Mesh::Normal(int faceIdx, int vertIdx) const
{
    return Normalize(m_Model.GetNormal(m_FaceNormalIndices[faceIdx * 3 + vertIdx]));
}
Vector3f Mesh::Tangent(int faceIdx, int vertIdx) const
{
    return Normalize(m_Model.GetTangent(m_FaceTangentIndices[faceIdx * 3 + vertIdx]));
}

int Mesh::DeviceId() const
{
    if (m_DeviceId < 0) m_Model.UploadToDevice();
    return m_DeviceId;
}
