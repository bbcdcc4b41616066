// scene.h — Scene of the drop-in facade (reference src/scene.h:13-72): the parsed .scene file.
#pragma once

#include <cassert>
#include <memory>
#include <string>
#include <vector>

#include "camera.h"
#include "light.h"
#include "model.h"

class Scene
{
public:
    explicit Scene(const std::string& filename);

    int   GetWidth() const { return m_Width; }
    int   GetHeight() const { return m_Height; }
    Float GetRatio() const { return Float(m_Width) / m_Height; }

    bool IsSSAAOn() const { return m_SSAA; }
    int  GetSSAAKernelSize() const { return m_SSAAKernelSize; }
    bool IsSSAOOn() const { return m_SSAO; }

    const PointLight& GetPointLight() const { assert(m_PointLight); return *m_PointLight; }
    PointLight&       GetPointLight() { assert(m_PointLight); return *m_PointLight; }  // multi-frame use (animated light)
    const DirLight&   GetDirLight() const { assert(m_DirLight); return *m_DirLight; }
    const Camera&     GetCamera() const { assert(m_Camera); return *m_Camera; }
    Camera&           GetCamera() { assert(m_Camera); return *m_Camera; }
    Camera::ProjectionType GetProjectionType() const { return m_ProjectionType; }

    unsigned int GetModelCount() const { return (unsigned int)m_Models.size(); }
    const Model& GetModel(int index) const { return *m_Models[index]; }
    Matrix4x4f   GetModelMatrix(int index) const { return m_ModelMatrices[index]; }
    bool         IsValid() const { return m_Valid; }

private:
    int  m_Width = 1280, m_Height = 800;
    bool m_SSAA = false;
    int  m_SSAAKernelSize = 2;
    bool m_SSAO = false;
    bool m_Valid = true;
    std::unique_ptr<PointLight> m_PointLight;
    std::unique_ptr<DirLight>   m_DirLight;
    std::unique_ptr<Camera>     m_Camera;
    Camera::ProjectionType      m_ProjectionType = Camera::Perspective;
    std::vector<std::unique_ptr<Model>> m_Models;
    std::vector<Matrix4x4f>             m_ModelMatrices;
};
