// material.h — Material / PBRMaterial containers (reference src/materials/material.h:13-41,
// pbrmaterial.h:13-46): same field names, feed the device material table.
#pragma once

#include <memory>
#include <string>

#include "geometry.h"
#include "texture.h"

class Material
{
public:
    std::string name;
    explicit Material(std::string n = "") : name(std::move(n)), ka(0.f), kd(0.f), ks(0.f), ke(0.f) {}

    Vector3f ka, kd, ks, ke;
    std::shared_ptr<Texture> diffuseMap, specularMap, normalMap, emissiveMap;

    bool HasDiffuseMap() const { return diffuseMap != nullptr; }
    bool HasSpecularMap() const { return specularMap != nullptr; }
    bool HasNormalMap() const { return normalMap != nullptr; }
    bool HasEmissiveMap() const { return emissiveMap != nullptr; }
};

class PBRMaterial
{
public:
    std::string name;
    explicit PBRMaterial(std::string n = "")
        : name(std::move(n)), ka(0.f), ke(0.f), albedo(1.f), roughness(0.f), metalness(0.f)
    {
    }

    Vector3f ka, ke, albedo;
    Float    roughness, metalness;
    std::shared_ptr<Texture> baseColorMap, roughnessMap, metalnessMap, ambientOcclusionMap, normalMap, emissiveMap;

    bool HasBaseColorMap() const { return baseColorMap != nullptr; }
    bool HasRoughnessMap() const { return roughnessMap != nullptr; }
    bool HasMetalnessMap() const { return metalnessMap != nullptr; }
    bool HasAmbientOcclusionMap() const { return ambientOcclusionMap != nullptr; }
    bool HasNormalMap() const { return normalMap != nullptr; }
    bool HasEmssiveMap() const { return emissiveMap != nullptr; }
};
