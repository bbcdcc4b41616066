// model.cpp — OBJ / MTL loading, load-time transforms and device upload.
//
// Written from scratch against the grammar and arithmetic of reference src/model.cpp:
//   Load order            model.cpp:16-73    (tangents BEFORE normalisation; PBR flag = last mesh in map order)
//   OBJ keywords          model.cpp:143-255  (mtllib, v, vt [3rd value ignored], vn, g, usemtl, f v/t/n fan)
//   MTL keywords          model.cpp:258-408  (newmtl Ka Kd Ks Ke Pr Pm map_Kd map_Ks map_Ke map_Bump|norm map_Ao map_Pr map_Pm)
//   texture load + v flip model.cpp:411-427  (wrap/filter captured from ForkerGL statics at load time)
//   vertex normalisation  model.cpp:107-138  (common scale, per-axis translation, applied as mat4 x vec4 Dot)
//   tangents              model.cpp:430-476  (accumulated over meshes in map order, faces in file order)
// Number parsing goes through std::istringstream like the reference so that odd tokens behave identically.
#include "model.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <charconv>
#include <cstring>
#include <fstream>
#include <limits>
#include <sstream>
#include <stdexcept>

#include "forkergl.h"
#include "forkergl_b200.h"
#include "shader.h"

static std::string LeftTrim(const std::string& s)
{
    size_t p = s.find_first_not_of(" \n\r\t\f\v");
    return p == std::string::npos ? "" : s.substr(p);
}
static bool StartsWith(const std::string& line, const char* key)
{
    return line.compare(0, strlen(key), key) == 0;
}

std::unique_ptr<Model> Model::Load(const std::string& filename, bool normalized, bool generateTangent,
                                   bool flipTexCoordY)
{
    std::unique_ptr<Model> model(new Model());
    model->m_HasTangents = generateTangent;
    if (!model->loadObjectFile(filename, flipTexCoordY))
    {
        fprintf(stderr, "[error] Failed to load model '%s'\n", filename.c_str());
        return nullptr;
    }
    if (generateTangent) model->generateTangents();
    if (normalized) model->normalizePositionVertices();

    // The model-level PBR switch is whatever the LAST mesh (alphabetical) says (reference model.cpp:52-68).
    for (auto& kv : model->m_Meshes)
    {
        std::shared_ptr<const PBRMaterial> pbr = kv.second->GetPBRMaterial();
        model->m_SupportPBR = pbr && (pbr->HasMetalnessMap() || pbr->HasRoughnessMap());
    }
    return model;
}

void Model::Render(Shader& shader) const
{
    for (auto& kv : m_Meshes) kv.second->Draw(shader);
}

int Model::GetNumFaces() const
{
    int total = 0;
    for (auto& kv : m_Meshes) total += kv.second->NumFaces();
    return total;
}

void Model::normalizePositionVertices()
{
    const Float MaxFloat = std::numeric_limits<Float>::max(), MinFloat = std::numeric_limits<Float>::min();
    Float xmin = MaxFloat, xmax = MinFloat, ymin = MaxFloat, ymax = MinFloat, zmin = MaxFloat, zmax = MinFloat;
    for (const Vector3f& v : m_Verts)
    {
        xmin = std::min(xmin, v.x), xmax = std::max(xmax, v.x);
        ymin = std::min(ymin, v.y), ymax = std::max(ymax, v.y);
        zmin = std::min(zmin, v.z), zmax = std::max(zmax, v.z);
    }
    Float      scaleFactor = 2.f / std::max(xmax - xmin, std::max(ymax - ymin, zmax - zmin));
    Matrix4x4f m(1.f);
    m[0][0] = scaleFactor;
    m[1][1] = scaleFactor;
    m[2][2] = scaleFactor;
    m[0][3] = -(xmax + xmin) / (xmax - xmin);
    m[1][3] = -(ymax + ymin) / (ymax - ymin);
    m[2][3] = -(zmax + zmin) / (zmax - zmin);
    for (Vector3f& v : m_Verts) v = (m * Vector4f(v, 1.f)).xyz();
}

// ---- OBJ reading -----------------------------------------------------------------------------------------------------
// The reference reads every line through std::istringstream (model.cpp:143-255), which takes minutes on the 10 M-triangle
// workload.  Here the file is mapped and scanned in place: "v" / "vt" / "vn" / "f" lines in plain form
// — decimal tokens separated by blanks, "a/b/c" corners — are decoded with std::from_chars, which like the stream
// extraction of the reference (libstdc++ num_get -> strtof) returns the correctly rounded float, so both produce the
// same bits.  Any line that is not in that plain form (a sign, hex, "1//3", junk behind a number ...) goes through the
// same istringstream statements as before, so odd files behave exactly as they did.
namespace
{
inline bool IsBlank(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\f' || c == '\v'; }

// one float token in plain form; advances p behind it
inline bool FastFloat(const char*& p, const char* end, float& out)
{
    while (p < end && IsBlank(*p)) ++p;
    const char* b = p;
    while (p < end && !IsBlank(*p)) ++p;
    if (b == p) return false;
    for (const char* q = b; q < p; ++q)
        if (!((*q >= '0' && *q <= '9') || *q == '-' || *q == '.' || *q == 'e' || *q == 'E')) return false;
    auto r = std::from_chars(b, p, out);
    if (r.ec != std::errc() || r.ptr != p) return false;
    float a = out < 0 ? -out : out;
    return a == 0.f || a >= std::numeric_limits<float>::min();  // subnormals / underflow: let the stream decide
}
inline bool FastUnsigned(const char*& p, const char* end, unsigned& out)
{
    const char* b = p;
    unsigned long long v = 0;
    while (p < end && *p >= '0' && *p <= '9' && p - b < 10) v = v * 10 + (unsigned)(*p++ - '0');
    if (b == p || v > 0xffffffffull || (p < end && *p >= '0' && *p <= '9')) return false;
    out = (unsigned)v;
    return true;
}
struct Corner { int v, t, n; };
// "f a/b/c a/b/c ..." in plain form
inline bool FastFace(const char* p, const char* end, std::vector<Corner>& corners)
{
    corners.clear();
    for (;;)
    {
        while (p < end && IsBlank(*p)) ++p;
        if (p == end) return true;
        unsigned v, t, n;
        if (!FastUnsigned(p, end, v) || p == end || *p++ != '/' || !FastUnsigned(p, end, t) || p == end || *p++ != '/' || !FastUnsigned(p, end, n))
            return false;
        if (p < end && !IsBlank(*p)) return false;
        corners.push_back({ (int)(v - 1), (int)(t - 1), (int)(n - 1) });
    }
}
}  // namespace

bool Model::loadObjectFile(const std::string& filename, bool flipVertically)
{
    int fd = open(filename.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat sb;
    if (fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode))
    {
        close(fd);
        return false;
    }
    const size_t size = (size_t)sb.st_size;
    const char*  data = size ? (const char*)mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0) : "";
    close(fd);
    if (size && data == (const char*)MAP_FAILED) return false;

    std::string         line, meshName, materialName;
    std::vector<Corner> corners;
    Mesh*               mesh = nullptr;  // m_Meshes[meshName], looked up once per "g" line instead of once per face
    auto addFace = [&]() {
        if (!mesh) mesh = m_Meshes[meshName].get();
        if (!mesh) throw std::runtime_error("OBJ: face before any 'g' line in " + filename + " (the reference dereferences a null mesh here)");
        for (size_t i = 1; i + 1 < corners.size(); ++i)  // triangle fan around corner 0
        {
            const Corner* tri[3] = { &corners[0], &corners[i], &corners[i + 1] };
            for (const Corner* c : tri)
            {
                mesh->AddVertIndex(c->v);
                mesh->AddTexCoordIndex(c->t);
                mesh->AddNormalIndex(c->n);
            }
        }
    };
    const char* const fileEnd = data + size;
    for (const char* cur = data; cur <= fileEnd;)
    {
        const char* nl = (const char*)memchr(cur, '\n', (size_t)(fileEnd - cur));
        const char* b = cur, *e = nl ? nl : fileEnd;
        cur = e + 1;  // (a file without a final newline ends the loop after its last line)
        while (b < e && (IsBlank(*b) || *b == '\n')) ++b;
        if (b == e) continue;
        // plain numeric lines, decoded in place
        if (b[0] == 'v' && e - b > 2)
        {
            const char* p = b + 2;
            if (b[1] == ' ')
            {
                Vector3f v;
                if (FastFloat(p, e, v.x) && FastFloat(p, e, v.y) && FastFloat(p, e, v.z)) { m_Verts.push_back(v); continue; }
            }
            else if (b[1] == 'n' && b[2] == ' ')
            {
                Vector3f n;
                ++p;
                if (FastFloat(p, e, n.x) && FastFloat(p, e, n.y) && FastFloat(p, e, n.z)) { m_Normals.push_back(n); continue; }
            }
            else if (b[1] == 't' && b[2] == ' ')
            {
                Vector2f t;
                ++p;
                if (FastFloat(p, e, t.x) && FastFloat(p, e, t.y)) { m_TexCoords.push_back(t); continue; }
            }
        }
        else if (b[0] == 'f' && e - b > 2 && b[1] == ' ' && FastFace(b + 2, e, corners))
        {
            addFace();
            continue;
        }
        // everything else: the statements of the reference's parser on this line
        line.assign(b, e);
        std::istringstream iss(line.c_str());
        char               ch;
        std::string        word;
        if (StartsWith(line, "mtllib "))
        {
            std::string mtl;
            iss >> word >> mtl;
            size_t      slash = filename.find_last_of("/");
            std::string dir = slash == std::string::npos ? "" : filename.substr(0, slash + 1);
            loadMaterials(dir, mtl, flipVertically);
        }
        else if (StartsWith(line, "v "))
        {
            Vector3f v;
            iss >> word >> v.x >> v.y >> v.z;
            m_Verts.push_back(v);
        }
        else if (StartsWith(line, "vt "))
        {
            Vector2f t;
            iss >> word >> t.x >> t.y;
            m_TexCoords.push_back(t);
        }
        else if (StartsWith(line, "vn "))
        {
            Vector3f n;
            iss >> word >> n.x >> n.y >> n.z;
            m_Normals.push_back(n);
        }
        else if (StartsWith(line, "g "))
        {
            iss >> ch >> meshName;
            m_Meshes[meshName] = std::make_shared<Mesh>(*this);
            mesh = m_Meshes[meshName].get();
        }
        else if (StartsWith(line, "usemtl "))
        {
            iss >> word >> materialName;
            m_Meshes[meshName]->SetMaterial(m_Materials[materialName]);
            m_Meshes[meshName]->SetPBRMaterial(m_PBRMaterials[materialName]);
        }
        else if (StartsWith(line, "f "))
        {
            iss >> ch;
            corners.clear();
            unsigned int v, t, n;
            while (iss >> v >> ch >> t >> ch >> n) corners.push_back({ (int)(v - 1), (int)(t - 1), (int)(n - 1) });
            addFace();
        }
    }
    if (size) munmap((void*)data, size);
    return true;
}

void Model::loadMaterials(const std::string& directory, const std::string& filename, bool flipVertically)
{
    std::ifstream in(directory + filename);
    if (in.fail())
    {
        fprintf(stderr, "[error] Cannot open the .mtl file: '%s'\n", (directory + filename).c_str());
        return;
    }
    std::string line, name;
    while (!in.eof())
    {
        std::getline(in, line);
        line = LeftTrim(line);
        std::istringstream iss(line.c_str());
        std::string        word, file;
        auto               vec3 = [&]() { Vector3f f; iss >> word >> f.x >> f.y >> f.z; return f; };
        auto               path = [&]() { iss >> word >> file; return directory + file; };

        if (StartsWith(line, "newmtl "))
        {
            iss >> word >> name;
            m_Materials[name] = std::make_shared<Material>(name);
            m_PBRMaterials[name] = std::make_shared<PBRMaterial>(name);
        }
        else if (StartsWith(line, "Ka ")) { Vector3f f = vec3(); m_Materials[name]->ka = f; m_PBRMaterials[name]->ka = f; }
        else if (StartsWith(line, "Kd ")) { Vector3f f = vec3(); m_Materials[name]->kd = f; m_PBRMaterials[name]->albedo = f; }
        else if (StartsWith(line, "Ks ")) { m_Materials[name]->ks = vec3(); }
        else if (StartsWith(line, "Ke ")) { Vector3f f = vec3(); m_Materials[name]->ke = f; m_PBRMaterials[name]->ke = f; }
        else if (StartsWith(line, "Pr ")) { float f; iss >> word >> f; m_PBRMaterials[name]->roughness = f; }
        else if (StartsWith(line, "Pm ")) { float f; iss >> word >> f; m_PBRMaterials[name]->metalness = f; }
        else if (StartsWith(line, "map_Kd "))
        {
            std::string p = path();
            loadTexture(p, m_Materials[name]->diffuseMap, flipVertically);
            loadTexture(p, m_PBRMaterials[name]->baseColorMap, flipVertically);
        }
        else if (StartsWith(line, "map_Ks ")) { loadTexture(path(), m_Materials[name]->specularMap, flipVertically); }
        else if (StartsWith(line, "map_Ke "))
        {
            std::string p = path();
            loadTexture(p, m_Materials[name]->emissiveMap, flipVertically);
            loadTexture(p, m_PBRMaterials[name]->emissiveMap, flipVertically);
        }
        else if (StartsWith(line, "map_Bump ") || StartsWith(line, "norm "))
        {
            std::string p = path();
            loadTexture(p, m_Materials[name]->normalMap, flipVertically);
            loadTexture(p, m_PBRMaterials[name]->normalMap, flipVertically);
        }
        else if (StartsWith(line, "map_Ao ")) { loadTexture(path(), m_PBRMaterials[name]->ambientOcclusionMap, flipVertically); }
        else if (StartsWith(line, "map_Pr ")) { loadTexture(path(), m_PBRMaterials[name]->roughnessMap, flipVertically); }
        else if (StartsWith(line, "map_Pm ")) { loadTexture(path(), m_PBRMaterials[name]->metalnessMap, flipVertically); }
    }
}

void Model::loadTexture(const std::string& textureFilename, std::shared_ptr<Texture>& texture, bool flipVertically)
{
    TGAImage image;
    if (!image.ReadTgaFile(textureFilename))
    {
        fprintf(stderr, "[warning] Failed to load texture: %s\n", textureFilename.c_str());
        return;
    }
    if (flipVertically) image.FlipVertically();
    texture = std::make_shared<Texture>(image, ForkerGL::TextureWrapping, ForkerGL::TextureFiltering);
}

void Model::generateTangents()
{
    m_Tangents.assign(m_Verts.size(), Vector3f(0.f));
    for (auto& kv : m_Meshes)
    {
        Mesh& mesh = *kv.second;
        for (int f = 0; f < mesh.NumFaces(); ++f)
        {
            int      i0 = mesh.GetVertIndex(f, 0), i1 = mesh.GetVertIndex(f, 1), i2 = mesh.GetVertIndex(f, 2);
            Vector3f e1 = mesh.Vert(f, 1) - mesh.Vert(f, 0);
            Vector3f e2 = mesh.Vert(f, 2) - mesh.Vert(f, 0);
            Vector2f d1 = mesh.TexCoord(f, 1) - mesh.TexCoord(f, 0);
            Vector2f d2 = mesh.TexCoord(f, 2) - mesh.TexCoord(f, 0);
            Float    det = d1.s * d2.t - d2.s * d1.t;
            if (det != 0.f)
            {
                Float    inv = 1.f / det;
                Vector3f T = Normalize(inv * Vector3f(d2.t * e1.x - d1.t * e2.x, d2.t * e1.y - d1.t * e2.y,
                                                      d2.t * e1.z - d1.t * e2.z));
                m_Tangents[i0] += T;
                m_Tangents[i1] += T;
                m_Tangents[i2] += T;
            }
            mesh.AddTangentIndex(i0);
            mesh.AddTangentIndex(i1);
            mesh.AddTangentIndex(i2);
        }
    }
    for (Vector3f& v : m_Tangents) v = (v.Length() == 0.f) ? Vector3f(1, 0, 0) : Normalize(v);
}

static void PutMap(int& dst, const std::shared_ptr<Texture>& t) { dst = t ? t->DeviceId() : -1; }
static void Put3(float* dst, const Vector3f& v) { dst[0] = v.x, dst[1] = v.y, dst[2] = v.z; }

void Model::UploadToDevice() const
{
    if (m_DeviceVertices >= 0) return;
    fgl_ctx* ctx = ForkerGL::Context();
    static_assert(sizeof(Vector3f) == 12 && sizeof(Vector2f) == 8, "vertex arrays are passed as packed floats");
    ForkerGL::Check(fgl_upload_vertices(ctx, m_Verts.empty() ? nullptr : &m_Verts[0].x, (int)m_Verts.size(),
                                        m_TexCoords.empty() ? nullptr : &m_TexCoords[0].x, (int)m_TexCoords.size(),
                                        m_Normals.empty() ? nullptr : &m_Normals[0].x, (int)m_Normals.size(),
                                        m_Tangents.empty() ? nullptr : &m_Tangents[0].x, (int)m_Tangents.size(),
                                        &m_DeviceVertices),
                    "upload vertices");
    for (auto& kv : m_Meshes)
    {
        const Mesh& mesh = *kv.second;
        FglMaterial m;
        memset(&m, 0, sizeof m);
        std::shared_ptr<const Material>    mat = mesh.GetMaterial();
        std::shared_ptr<const PBRMaterial> pbr = mesh.GetPBRMaterial();
        if (!mat || !pbr) throw std::runtime_error("Model: mesh '" + kv.first + "' has no material (usemtl missing)");
        Put3(m.ka, mat->ka), Put3(m.kd, mat->kd), Put3(m.ks, mat->ks), Put3(m.ke, mat->ke);
        Put3(m.pbr_ke, pbr->ke), Put3(m.albedo, pbr->albedo);
        m.roughness = pbr->roughness, m.metalness = pbr->metalness;
        PutMap(m.diffuse_map, mat->diffuseMap), PutMap(m.specular_map, mat->specularMap);
        PutMap(m.normal_map, mat->normalMap), PutMap(m.emissive_map, mat->emissiveMap);
        PutMap(m.base_color_map, pbr->baseColorMap), PutMap(m.roughness_map, pbr->roughnessMap);
        PutMap(m.metalness_map, pbr->metalnessMap), PutMap(m.ao_map, pbr->ambientOcclusionMap);
        PutMap(m.pbr_normal_map, pbr->normalMap), PutMap(m.pbr_emissive_map, pbr->emissiveMap);
        ForkerGL::Check(fgl_upload_mesh(ctx, m_DeviceVertices, mesh.NumFaces(), mesh.VertIndices().data(),
                                        mesh.TexCoordIndices().data(), mesh.NormalIndices().data(), &m,
                                        m_HasTangents ? 1 : 0, m_SupportPBR ? 1 : 0, &mesh.m_DeviceId),
                        "upload mesh");
    }
}
