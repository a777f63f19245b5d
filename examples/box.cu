// examples/box.cu -- the reference's src/examples/Box.cpp:36-202 on the B200 API, headless: the textured cube with
// perspective-correct UVs, the screen-space UV derivatives from PixelData::computePerspectiveDerivatives and the
// anisotropic / trilinear sampler of Texture.h (here swr/Texture.h, device code), RasterMode::Span, CullMode::CW,
// camera on the Box.cpp:190-195 orbit at 0, 0.5 and 2.0 rad.  The reference keeps the matrix and the texture in static
// shader members; nvcc has no memory-space qualifier for data members, so they travel in the uniform block
// (swr::uniforms<T>()).  The mesh, the three matrices and the texels come from examples/data/box_scene.bin
// (examples/data/make_box_scene.py).  Prints the fragment count per camera angle: the reference draws 37 574 / 39 799 /
// 39 227.  This file is compiled WITHOUT -fmad=false: the vertex shader's matrix product is
// spelled with the exact helpers, so contraction cannot move a vertex.
//   usage: box [path/to/box_scene.bin] [out.ppm]
#include <swr/Renderer.h>
#include <swr/Texture.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace swr;

struct Uniforms {
    float mvp[16];           // row-major
    TextureView texture;
};

struct VertexArrayData {     // ObjData::VertexArrayData: vertex, normal, texcoord
    float vertex[3], normal[3], texcoord[2];
};

class PixelShader : public PixelShaderBase<PixelShader> {
public:
    static const bool InterpolateZ = false;
    static const bool InterpolateW = true;  // Required for perspective correct texturing
    static const int AVarCount = 0;
    static const int PVarCount = 2;         // UV coordinates
    static const int RenderTargets = 1;

    __device__ static void drawPixel(const PixelData &p)
    {
        // Compute texture coordinate derivatives
        float dudx, dudy, dvdx, dvdy;
        p.computePerspectiveDerivatives(*p.equations, 0, dudx, dudy); // U derivatives
        p.computePerspectiveDerivatives(*p.equations, 1, dvdx, dvdy); // V derivatives
        target<unsigned>(p, 0) = 0xff000000u | textureSample(uniforms<Uniforms>().texture, p.pvar[0], p.pvar[1], dudx, dvdx, dudy, dvdy);
    }
};

class VertexShader : public VertexShaderBase<VertexShader> {
public:
    static const int AttribCount = 1;
    static const int AVarCount = 0;
    static const int PVarCount = 2;

    __device__ static void processVertex(VertexShaderInput in, VertexShaderOutput *out)
    {
        using namespace swr::detail;
        const VertexArrayData *data = static_cast<const VertexArrayData *>(in[0]);
        const float *m = uniforms<Uniforms>().mvp;
        const float x = data->vertex[0], y = data->vertex[1], z = data->vertex[2];
        // position = modelViewProjectionMatrix * vec4f(vertex, 1), each row left to right, never contracted
        out->x = fadd(fadd(fadd(fmul(m[0], x), fmul(m[1], y)), fmul(m[2], z)), m[3]);
        out->y = fadd(fadd(fadd(fmul(m[4], x), fmul(m[5], y)), fmul(m[6], z)), m[7]);
        out->z = fadd(fadd(fadd(fmul(m[8], x), fmul(m[9], y)), fmul(m[10], z)), m[11]);
        out->w = fadd(fadd(fadd(fmul(m[12], x), fmul(m[13], y)), fmul(m[14], z)), m[15]);
        out->pvar[0] = data->texcoord[0];
        out->pvar[1] = data->texcoord[1];
    }
};

int main(int argc, char *argv[])
{
    std::string path = argc > 1 ? argv[1] : "";
    if (path.empty()) {
        std::string self = argv[0];
        const size_t cut = self.rfind('/');
        path = (cut == std::string::npos ? std::string(".") : self.substr(0, cut)) + "/../data/box_scene.bin";
    }
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) { std::fprintf(stderr, "cannot open %s\n", path.c_str()); return 2; }
    int hdr[5];
    if (std::fread(hdr, sizeof(int), 5, f) != 5 || hdr[0] != 0x584F4253) { std::fprintf(stderr, "bad scene file\n"); return 2; }
    std::vector<VertexArrayData> vdata(hdr[1]);
    std::vector<int> idata(hdr[2]);
    float mvps[3][16];
    std::vector<uint32_t> texels((size_t)hdr[3] * hdr[4]);
    bool ok = std::fread(vdata.data(), sizeof(VertexArrayData), vdata.size(), f) == vdata.size();
    ok = ok && std::fread(idata.data(), sizeof(int), idata.size(), f) == idata.size();
    ok = ok && std::fread(mvps, sizeof(float), 48, f) == 48;
    ok = ok && std::fread(texels.data(), 4, texels.size(), f) == texels.size();
    std::fclose(f);
    if (!ok) { std::fprintf(stderr, "short scene file\n"); return 2; }

    Rasterizer r;
    VertexProcessor v(&r);

    r.setRasterMode(RasterMode::Span);
    r.setScissorRect(0, 0, 640, 480);
    r.setPixelShader<PixelShader>();

    v.setViewport(0, 0, 640, 480);
    v.setCullMode(CullMode::CW);
    v.setVertexShader<VertexShader>();

    unsigned *screen = static_cast<unsigned *>(swr_device_alloc(r.context(), sizeof(unsigned) * 640 * 480));
    r.setRenderTarget(0, screen, 640 * 4, 640, 480);

    // Texture(surface): the mip chain (Texture.h:220-293), in device memory
    std::vector<uint32_t> chain(texels.size() * 2 + 16);
    int32_t mw[kMaxMipLevels], mh[kMaxMipLevels];
    int64_t moff[kMaxMipLevels];
    Uniforms u;
    std::memset(&u, 0, sizeof(u));
    u.texture.levels = buildMipChain(texels.data(), hdr[3], hdr[4], chain.data(), mw, mh, moff);
    uint32_t *dchain = static_cast<uint32_t *>(swr_device_alloc(r.context(), chain.size() * 4));
    swr_memcpy_h2d(r.context(), dchain, chain.data(), chain.size() * 4);
    for (int i = 0; i < u.texture.levels; ++i) { u.texture.level[i] = dchain + moff[i]; u.texture.w[i] = mw[i]; u.texture.h[i] = mh[i]; }
    u.texture.maxAnisotropy = 8;

    std::vector<unsigned> host(640 * 480);
    for (int frame = 0; frame < 3; ++frame) {
        std::memcpy(u.mvp, mvps[frame], sizeof(u.mvp));
        r.setUniforms(&u, sizeof(u));
        swr_memset32(r.context(), screen, 0, 640 * 480);            // SDL_FillRect(screen, NULL, 0)
        swr_reset_stats(r.context());

        // Draw the box
        v.setVertexAttribPointer(0, sizeof(VertexArrayData), &vdata[0]);
        v.drawElements(DrawMode::Triangle, idata.size(), &idata[0]);

        swr_stats st;
        swr_get_stats(r.context(), &st);
        swr_memcpy_d2h(r.context(), host.data(), screen, sizeof(unsigned) * host.size());
        r.finish();
        long covered = 0;
        unsigned long long sum = 0;
        for (unsigned px : host) { covered += px != 0; sum += px & 0xffffffu; }
        std::printf("frame %d: fragments %llu covered %ld checksum %llu\n", frame, (unsigned long long)st.fragments, covered, sum);
    }
    if (argc > 2) {
        FILE *o = std::fopen(argv[2], "wb");
        if (o) {
            std::fprintf(o, "P6\n640 480\n255\n");
            for (unsigned px : host) { unsigned char c[3] = { (unsigned char)(px >> 16), (unsigned char)(px >> 8), (unsigned char)px }; std::fwrite(c, 1, 3, o); }
            std::fclose(o);
        }
    }
    swr_device_free(r.context(), dchain);
    swr_device_free(r.context(), screen);
    return 0;
}
