// swr/Texture.h -- the reference's texture sampler as device code.
//
// Reference: src/examples/Texture.h -- wrap (:41-44), anisotropic footprint from the UV derivatives
// (:52-69), major-axis sample loop (:78-117), trilinear (:123-143), bilinear in uint8 math (:145-218), mip
// chain by 2x2 box filter (:220-293).  Same expressions in the same order; the float -> Uint8 casts go
// through int, which is what the reference compiles to on x86 (negative products in lerpColors wrap
// instead of saturating).  The mip chain is built on the host by the same integer box filter
// (softwarerenderer_b200.api.build_mip_chain / swr::buildMipChain) and handed over as a TextureView.
//
// fmodf / floorf / ceilf / sqrtf are exact or correctly rounded on both sides; log2f is not specified
// to the last bit, so against a glibc build the level-of-detail fraction can differ in its last ulp and
// a blended channel by 1 LSB on rare pixels (tests allow exactly that).
#pragma once

#include <math.h>
#include <stdint.h>
#include "detail/common.h"

namespace swr {

constexpr int kMaxMipLevels = 14;

/// Mip chain of 0x00RRGGBB texels in device memory; level 0 is the base texture.
struct TextureView {
    const uint32_t *level[kMaxMipLevels];
    int32_t w[kMaxMipLevels];
    int32_t h[kMaxMipLevels];
    int32_t levels;
    int32_t maxAnisotropy;       // Texture.h:14-16: clamp(maxAnisotropy, 1, 16), default 8
};

namespace detail {
SWR_HD int texR(uint32_t c) { return (int)((c >> 16) & 0xFF); }
SWR_HD int texG(uint32_t c) { return (int)((c >> 8) & 0xFF); }
SWR_HD int texB(uint32_t c) { return (int)(c & 0xFF); }
SWR_HD uint32_t texPack(int r, int g, int b) { return ((uint32_t)(r & 0xFF) << 16) | ((uint32_t)(g & 0xFF) << 8) | (uint32_t)(b & 0xFF); }
SWR_HD int texU8(float v) { return (int)v & 0xFF; }                          // (Uint8)float as x86 does it

// Texture.h:191-196
SWR_HD uint32_t texLerp(uint32_t c1, uint32_t c2, float t)
{
    const int r = texR(c1) + texU8((float)(texR(c2) - texR(c1)) * t);
    const int g = texG(c1) + texU8((float)(texG(c2) - texG(c1)) * t);
    const int b = texB(c1) + texU8((float)(texB(c2) - texB(c1)) * t);
    return texPack(r, g, b);
}

SWR_HD int texBilerpChannel(int c00, int c10, int c01, int c11, float fx, float fy)
{
    // Texture.h:199-204: c00*(1-fx)*(1-fy) + c10*fx*(1-fy) + c01*(1-fx)*fy + c11*fx*fy, left to right
    const float v = (float)c00 * (1 - fx) * (1 - fy) + (float)c10 * fx * (1 - fy) + (float)c01 * (1 - fx) * fy + (float)c11 * fx * fy;
    return texU8(v);
}

// Texture.h:145-180
SWR_HD uint32_t texBilinear(const TextureView &t, int mip, float u, float v)
{
    if (mip < 0 || mip >= t.levels) return 0;
    const int w = t.w[mip], h = t.h[mip];
    const uint32_t *px = t.level[mip];
    const float fpx = u * (float)(w - 1);
    const float fpy = v * (float)(h - 1);
    int x0 = (int)floorf(fpx); if (x0 < 0) x0 = 0;
    int y0 = (int)floorf(fpy); if (y0 < 0) y0 = 0;
    const int x1 = x0 + 1 < w - 1 ? x0 + 1 : w - 1;
    const int y1 = y0 + 1 < h - 1 ? y0 + 1 : h - 1;
    const float fx = fpx - (float)x0;
    const float fy = fpy - (float)y0;
    const uint32_t c00 = px[y0 * w + x0], c10 = px[y0 * w + x1], c01 = px[y1 * w + x0], c11 = px[y1 * w + x1];
    return texPack(texBilerpChannel(texR(c00), texR(c10), texR(c01), texR(c11), fx, fy),
                   texBilerpChannel(texG(c00), texG(c10), texG(c01), texG(c11), fx, fy),
                   texBilerpChannel(texB(c00), texB(c10), texB(c01), texB(c11), fx, fy));
}

// Texture.h:123-143
SWR_HD uint32_t texTrilinear(const TextureView &t, float u, float v, float rho)
{
    if (t.levels <= 0) return 0;
    float lod = log2f(rho > 1e-6f ? rho : 1e-6f);
    const float top = (float)(t.levels - 1);
    lod = lod < 0.0f ? 0.0f : (top < lod ? top : lod);                       // std::clamp
    const int lodBase = (int)floorf(lod);
    const int lodNext = lodBase + 1 < t.levels - 1 ? lodBase + 1 : t.levels - 1;
    float lodFrac = lod - (float)lodBase;
    lodFrac = lodFrac < 0.0f ? 0.0f : (1.0f < lodFrac ? 1.0f : lodFrac);
    const uint32_t cBase = texBilinear(t, lodBase, u, v);
    if (lodBase != lodNext) return texLerp(cBase, texBilinear(t, lodNext, u, v), lodFrac);
    return cBase;
}

SWR_HD float texWrap(float x)
{
    x = fmodf(x, 1.0f);
    if (x < 0) x += 1.0f;
    return x;
}
} // namespace detail

/// Texture::sample (Texture.h:35-118).
SWR_HD uint32_t textureSample(const TextureView &t, float u, float v, float dudx, float dvdx, float dudy, float dvdy)
{
    using namespace detail;
    if (t.levels <= 0) return 0;
    u = texWrap(u);
    v = texWrap(v);
    const float dudx_s = dudx * (float)t.w[0], dvdx_s = dvdx * (float)t.h[0];
    const float dudy_s = dudy * (float)t.w[0], dvdy_s = dvdy * (float)t.h[0];
    float dx_len = sqrtf(dudx_s * dudx_s + dvdx_s * dvdx_s);
    float dy_len = sqrtf(dudy_s * dudy_s + dvdy_s * dvdy_s);
    dx_len = dx_len < 1e-6f ? 1e-6f : dx_len;                                // std::max(x, 1e-6f)
    dy_len = dy_len < 1e-6f ? 1e-6f : dy_len;
    const float major_len = dx_len < dy_len ? dy_len : dx_len;
    const float minor_len = dy_len < dx_len ? dy_len : dx_len;
    const float maxAniso = (float)t.maxAnisotropy;
    float ratio = major_len / minor_len;
    ratio = maxAniso < ratio ? maxAniso : ratio;
    int num_samples = (int)ceilf(ratio);
    if (num_samples < 1) num_samples = 1;
    if (num_samples <= 1) return texTrilinear(t, u, v, major_len);

    float major_du, major_dv;
    if (dx_len > dy_len) { major_du = dudx / dx_len; major_dv = dvdx / dx_len; }
    else                 { major_du = dudy / dy_len; major_dv = dvdy / dy_len; }
    float r = 0, g = 0, b = 0;
    const float step = 1.0f / (float)num_samples;
    for (int i = 0; i < num_samples; ++i) {
        const float tt = ((float)i + 0.5f) * step - 0.5f;
        const float su = texWrap(u + major_du * major_len * tt);
        const float sv = texWrap(v + major_dv * major_len * tt);
        const uint32_t c = texTrilinear(t, su, sv, minor_len);
        r += (float)texR(c);
        g += (float)texG(c);
        b += (float)texB(c);
    }
    r = r / (float)num_samples; r = 255.0f < r ? 255.0f : r;
    g = g / (float)num_samples; g = 255.0f < g ? 255.0f : g;
    b = b / (float)num_samples; b = 255.0f < b ? 255.0f : b;
    return texPack(texU8(r), texU8(g), texU8(b));
}

/// Host side: the mip chain of Texture::generateMipmaps (Texture.h:220-293), appended level after level
/// to `out` (level 0 first); returns the number of levels and fills w[] / h[] / offset[] (in texels).
inline int buildMipChain(const uint32_t *base, int width, int height, uint32_t *out, int32_t *w, int32_t *h, int64_t *offset)
{
    int levels = 0;
    int64_t pos = 0;
    for (int i = 0; i < width * height; ++i) out[i] = base[i] & 0x00FFFFFFu;
    w[0] = width; h[0] = height; offset[0] = 0;
    levels = 1;
    pos = (int64_t)width * height;
    while ((width > 1 || height > 1) && levels < kMaxMipLevels) {
        const int nw = width / 2 > 1 ? width / 2 : 1, nh = height / 2 > 1 ? height / 2 : 1;
        const uint32_t *src = out + offset[levels - 1];
        uint32_t *dst = out + pos;
        for (int y = 0; y < nh; ++y)
            for (int x = 0; x < nw; ++x) {
                const uint32_t p00 = src[(y * 2) * width + (x * 2)];
                const uint32_t p10 = x * 2 + 1 < width ? src[(y * 2) * width + (x * 2 + 1)] : p00;
                const uint32_t p01 = y * 2 + 1 < height ? src[(y * 2 + 1) * width + (x * 2)] : p00;
                const uint32_t p11 = (x * 2 + 1 < width && y * 2 + 1 < height) ? src[(y * 2 + 1) * width + (x * 2 + 1)] : p00;
                const int r = (detail::texR(p00) + detail::texR(p10) + detail::texR(p01) + detail::texR(p11)) >> 2;
                const int g = (detail::texG(p00) + detail::texG(p10) + detail::texG(p01) + detail::texG(p11)) >> 2;
                const int b = (detail::texB(p00) + detail::texB(p10) + detail::texB(p01) + detail::texB(p11)) >> 2;
                dst[y * nw + x] = detail::texPack(r, g, b);
            }
        w[levels] = nw; h[levels] = nh; offset[levels] = pos;
        pos += (int64_t)nw * nh;
        width = nw; height = nh;
        ++levels;
    }
    return levels;
}

} // namespace swr
