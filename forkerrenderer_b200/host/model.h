// model.h — Model of the drop-in facade (reference src/model.h:15-67).  OBJ/MTL loading is load-time work and
// stays on the host, but its arithmetic feeds coverage, so the load-time transforms (vertex normalisation,
// tangent accumulation order, fan triangulation, texture flip) reproduce the reference's results bit for bit
// (SURVEY.md §7.2 last bullet).  Render(shader) submits the meshes in std::map (alphabetical) order, which
// defines the primitive ids.
#pragma once

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "geometry.h"
#include "material.h"
#include "mesh.h"

struct Shader;

class Model
{
public:
    static std::unique_ptr<Model> Load(const std::string& filename, bool normalized = false,
                                       bool generateTangent = false, bool flipTexCoordY = true);
    Model() = default;
    Model(const Model&) = delete;

    void Render(Shader& shader) const;

    Vector3f GetVert(int i) const { return m_Verts[i]; }
    Vector2f GetTexCoord(int i) const { return m_TexCoords[i]; }
    Vector3f GetNormal(int i) const { return m_Normals[i]; }
    Vector3f GetTangent(int i) const { return m_Tangents[i]; }
    int      GetNumVerts() const { return (int)m_Verts.size(); }
    int      GetNumFaces() const;
    bool     HasTangents() const { return m_HasTangents; }
    bool     SupportPBR() const { return m_SupportPBR; }

    const std::map<std::string, std::shared_ptr<Mesh>>& Meshes() const { return m_Meshes; }
    void UploadToDevice() const;  // idempotent

private:
    std::map<std::string, std::shared_ptr<Mesh>>        m_Meshes;
    std::map<std::string, std::shared_ptr<Material>>    m_Materials;
    std::map<std::string, std::shared_ptr<PBRMaterial>> m_PBRMaterials;
    std::vector<Vector3f> m_Verts, m_Normals, m_Tangents;
    std::vector<Vector2f> m_TexCoords;
    bool                  m_HasTangents = false;
    bool                  m_SupportPBR = false;
    mutable int           m_DeviceVertices = -1;

    bool loadObjectFile(const std::string& filename, bool flipVertically);
    void loadMaterials(const std::string& directory, const std::string& filename, bool flipVertically);
    void loadTexture(const std::string& textureFilename, std::shared_ptr<Texture>& texture, bool flipVertically);
    void normalizePositionVertices();
    void generateTangents();
};
