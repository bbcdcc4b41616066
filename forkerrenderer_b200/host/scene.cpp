// scene.cpp — .scene parser: the 8 keywords of reference src/scene.cpp:54-171
// (mode / screen / ssaa / ssao / shadow / light / camera / model, '#' comments), with the same side effects on
// ForkerGL's render mode and the shadow status.  A model that fails to load is reported and skipped (the
// reference pushes the nullptr and crashes later).
#include "scene.h"

#include <cstring>
#include <fstream>
#include <sstream>

#include "forkergl.h"
#include "shadow.h"

Scene::Scene(const std::string& filename)
{
    std::ifstream in(filename);
    if (in.fail())
    {
        fprintf(stderr, "[error] Failed to open the .scene file '%s'\n", filename.c_str());
        m_Valid = false;
        return;
    }
    std::string line;
    while (!in.eof())
    {
        std::getline(in, line);
        size_t p = line.find_first_not_of(" \n\r\t\f\v");
        line = p == std::string::npos ? "" : line.substr(p);
        std::istringstream iss(line.c_str());
        std::string        key, word;
        auto is = [&](const char* k) { return line.compare(0, strlen(k), k) == 0; };

        if (is("#")) continue;
        if (is("mode "))
        {
            iss >> key >> word;
            ForkerGL::SetRenderMode(word == "deferred" ? ForkerGL::Deferred : ForkerGL::Forward);
        }
        else if (is("screen ")) iss >> key >> m_Width >> m_Height;
        else if (is("ssaa "))
        {
            iss >> key >> word >> m_SSAAKernelSize;
            m_SSAA = (word == "on");
        }
        else if (is("ssao "))
        {
            iss >> key >> word;
            m_SSAO = (word == "on");
        }
        else if (is("shadow "))
        {
            iss >> key >> word;
            Shadow::SetShadowStatus(word == "on");
        }
        else if (is("light "))
        {
            iss >> key >> word;
            Vector3f a, c;
            if (word == "point")
            {
                iss >> a.x >> a.y >> a.z >> c.x >> c.y >> c.z;
                m_PointLight.reset(new PointLight(a, c));
            }
            else if (word == "dir")
            {
                iss >> a.x >> a.y >> a.z >> c.x >> c.y >> c.z;
                m_DirLight.reset(new DirLight(a, c));
            }
            else
                fprintf(stderr, "[warning] Invalid light type: %s\n", word.c_str());
        }
        else if (is("camera "))
        {
            iss >> key >> word;
            if (word == "persp") m_ProjectionType = Camera::Perspective;
            else if (word == "ortho") m_ProjectionType = Camera::Orthographic;
            else
            {
                fprintf(stderr, "[warning] Invalid camera type: %s\n", word.c_str());
                continue;
            }
            Point3f eye, at;
            iss >> eye.x >> eye.y >> eye.z >> at.x >> at.y >> at.z;
            m_Camera.reset(new Camera(eye, at));
        }
        else if (is("model "))
        {
            std::string file, b1, b2;
            Point3f     pos;
            Float       rotateY, scale;
            iss >> key >> file >> b1 >> b2 >> pos.x >> pos.y >> pos.z >> rotateY >> scale;
            std::unique_ptr<Model> m = Model::Load(file, b1 == "true", b2 == "true");
            if (!m)
            {
                m_Valid = false;
                continue;
            }
            m_Models.push_back(std::move(m));
            m_ModelMatrices.push_back(MakeModelMatrix(pos, rotateY, scale));
        }
    }
}
