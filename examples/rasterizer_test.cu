// examples/rasterizer_test.cu -- the reference's src/examples/RasterizerTest.cpp:37-81 on the B200 API: one
// screen-space triangle with RGB affine variables through Rasterizer::drawTriangle (no VertexProcessor), plus the
// single-primitive drawLine / drawPoint entries of Rasterizer.h:99-114.  The SDL window is replaced by a device
// frame buffer registered as render target 0; the program prints the fragment and covered-pixel counts (the
// reference draws 25 900 fragments in Span and 26 100 in Block mode).
//   usage: rasterizer_test [block]
#include <swr/Renderer.h>

#include <cstdio>
#include <vector>

using namespace swr;

class PixelShader : public PixelShaderBase<PixelShader> {
public:
    static const bool InterpolateZ = false;
    static const bool InterpolateW = false;
    static const int AVarCount = 3;
    static const int RenderTargets = 1;      // additive: slot 0 is staged per tile

    __device__ static void drawPixel(const PixelData &p)
    {
        int rint = (int)(p.avar[0] * 255);
        int gint = (int)(p.avar[1] * 255);
        int bint = (int)(p.avar[2] * 255);
        target<unsigned>(p, 0) = 0xff000000u | (unsigned)(rint << 16 | gint << 8 | bint);
    }
};

static void report(Rasterizer &r, unsigned *buffer, const char *what)
{
    swr_stats st;
    swr_get_stats(r.context(), &st);
    std::vector<unsigned> host(640 * 480);
    swr_memcpy_d2h(r.context(), host.data(), buffer, sizeof(unsigned) * host.size());
    r.finish();
    long covered = 0;
    for (unsigned px : host) covered += px != 0;
    std::printf("%s: fragments %llu covered %ld\n", what, (unsigned long long)st.fragments, covered);
}

int main(int argc, char *argv[])
{
    Rasterizer r;
    r.setScissorRect(0, 0, 640, 480);
    r.setPixelShader<PixelShader>();
    if (argc > 1 && argv[1][0] == 'b') r.setRasterMode(RasterMode::Block);

    unsigned *buffer = static_cast<unsigned *>(swr_device_alloc(r.context(), sizeof(unsigned) * 640 * 480));
    swr_memset32(r.context(), buffer, 0, 640 * 480);
    r.setRenderTarget(0, buffer, 640 * 4, 640, 480);

    RasterizerVertex v0 = {}, v1 = {}, v2 = {};
    v0.x = 320; v0.y = 100; v0.avar[0] = 1.0f; v0.avar[1] = 0.0f; v0.avar[2] = 0.0f;
    v1.x = 480; v1.y = 200; v1.avar[0] = 0.0f; v1.avar[1] = 1.0f; v1.avar[2] = 0.0f;
    v2.x = 120; v2.y = 300; v2.avar[0] = 0.0f; v2.avar[1] = 0.0f; v2.avar[2] = 1.0f;

    r.drawTriangle(v0, v1, v2);
    r.finish();
    report(r, buffer, "triangle");

    swr_reset_stats(r.context());
    swr_memset32(r.context(), buffer, 0, 640 * 480);
    r.drawLine(v0, v1);                      // DDA: max(|160|, |100|) = 160 steps (Rasterizer.h:175-194)
    r.drawPoint(v2);
    r.finish();
    report(r, buffer, "line+point");
    swr_device_free(r.context(), buffer);
    return 0;
}
