// examples/vertex_processor_test.cu -- the reference's src/examples/VertexProcessorTest.cpp:30-121 on the B200 API:
// one clip-space triangle that sticks out of the +-X planes, an offset viewport inside a larger scissor,
// CullMode::None.  Host vertex and index arrays are handed over exactly as the reference program does it
// (setVertexAttribPointer without a size, VertexProcessorTest.cpp:119-120).  Prints the fragment count: the
// reference draws 41 068 fragments in every raster mode.
//   usage: vertex_processor_test [span|block|adaptive]
#include <swr/Renderer.h>

#include <cstdio>
#include <vector>

using namespace swr;

class PixelShader : public PixelShaderBase<PixelShader> {
public:
    static const bool InterpolateZ = false;
    static const bool InterpolateW = false;
    static const int AVarCount = 3;
    static const int PVarCount = 0;
    static const int RenderTargets = 1;

    __device__ static void drawPixel(const PixelData &p)
    {
        int rint = (int)(p.avar[0] * 255);
        int gint = (int)(p.avar[1] * 255);
        int bint = (int)(p.avar[2] * 255);
        target<unsigned>(p, 0) = 0xff000000u | (unsigned)(rint << 16 | gint << 8 | bint);
    }
};

struct VertexData {
    float x, y, z;
    float r, g, b;
};

class VertexShader : public VertexShaderBase<VertexShader> {
public:
    static const int AttribCount = 1;
    static const int AVarCount = 3;
    static const int PVarCount = 0;

    __device__ static void processVertex(VertexShaderInput in, VertexShaderOutput *out)
    {
        const VertexData *data = static_cast<const VertexData *>(in[0]);
        out->x = data->x;
        out->y = data->y;
        out->z = data->z;
        out->w = 1.0f;
        out->avar[0] = data->r;
        out->avar[1] = data->g;
        out->avar[2] = data->b;
    }
};

int main(int argc, char *argv[])
{
    Rasterizer r;
    VertexProcessor v(&r);

    r.setScissorRect(0, 0, 640, 480);
    r.setPixelShader<PixelShader>();
    if (argc > 1 && argv[1][0] == 'b') r.setRasterMode(RasterMode::Block);
    if (argc > 1 && argv[1][0] == 'a') r.setRasterMode(RasterMode::Adaptive);

    unsigned *buffer = static_cast<unsigned *>(swr_device_alloc(r.context(), sizeof(unsigned) * 640 * 480));
    swr_memset32(r.context(), buffer, 0, 640 * 480);
    r.setRenderTarget(0, buffer, 640 * 4, 640, 480);

    v.setViewport(100, 100, 640 - 200, 480 - 200);
    v.setCullMode(CullMode::None);
    v.setVertexShader<VertexShader>();

    VertexData vdata[3];
    vdata[0] = { 0.0f, 0.5f, 0.0f, 1.0f, 0.0f, 0.0f };
    vdata[1] = { -1.5f, -0.5f, 0.0f, 0.0f, 1.0f, 0.0f };
    vdata[2] = { 1.5f, -0.5f, 0.0f, 0.0f, 0.0f, 1.0f };
    int idata[3] = { 0, 1, 2 };

    v.setVertexAttribPointer(0, sizeof(VertexData), vdata);
    v.drawElements(DrawMode::Triangle, 3, idata);

    swr_stats st;
    swr_get_stats(r.context(), &st);
    std::vector<unsigned> host(640 * 480);
    swr_memcpy_d2h(r.context(), host.data(), buffer, sizeof(unsigned) * host.size());
    r.finish();
    long covered = 0;
    for (unsigned px : host) covered += px != 0;
    std::printf("fragments %llu covered %ld\n", (unsigned long long)st.fragments, covered);
    swr_device_free(r.context(), buffer);
    return 0;
}
