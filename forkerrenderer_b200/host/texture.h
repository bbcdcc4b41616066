// texture.h — Texture of the drop-in facade (reference src/texture.h:11-145).  Owns a copy of the TGAImage and
// the wrap/filter modes captured at construction; sampling itself runs on the device (csrc/texture.cuh), so this
// class is a resource handle: DeviceId() uploads on first use.
#pragma once

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

    int DeviceId() const;  // fgl texture handle (uploads on first call)

private:
    int         m_Width, m_Height;
    TGAImage    m_Image;
    WrapMode    m_WrapMode;
    FilterMode  m_FilterMode;
    mutable int m_DeviceId = -1;
};
