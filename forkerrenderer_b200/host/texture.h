// texture.h — Texture of the drop-in facade (reference src/texture.h:11-145).  Owns a copy of the TGAImage and
// the wrap/filter modes captured at construction.  The device passes sample through csrc/device_math.cuh (tex_sample),
// DeviceId() uploads on first use; Sample / SampleFloat are the same arithmetic on the host for callers of the interface.
#pragma once

#include "geometry.h"
#include "tgaimage.h"

class Texture
{
public:
    enum WrapMode { NoWrap, Repeat, MirroredRepeat, ClampToEdge };
    enum FilterMode { Nearest, Linear };

    Texture(const TGAImage& img, WrapMode wrap = NoWrap, FilterMode filter = Nearest)
        : m_Width(img.GetWidth()), m_Height(img.GetHeight()), m_Image(img), m_WrapMode(wrap), m_FilterMode(filter)
    {
    }

    int             GetWidth() const { return m_Width; }
    int             GetHeight() const { return m_Height; }
    WrapMode        GetWrapMode() const { return m_WrapMode; }
    FilterMode      GetFilterMode() const { return m_FilterMode; }
    const TGAImage& GetImage() const { return m_Image; }

    Color3 Sample(const Vector2f& coord) const;       // reference texture.h:41-45
    Float  SampleFloat(const Vector2f& coord) const;  // reference texture.h:47-51

    int DeviceId() const;  // fgl texture handle (uploads on first call)

private:
    Vector3f filtered(Float u, Float v) const;  // wrap + filter, texel values in [0, 255]

    int         m_Width, m_Height;
    TGAImage    m_Image;
    WrapMode    m_WrapMode;
    FilterMode  m_FilterMode;
    mutable int m_DeviceId = -1;
};
