// examples/foreign_rasterizer.cu -- a VertexProcessor in front of a user-supplied IRasterizer (the reference's plug
// point, src/renderer/IRasterizer.h:56-71 / VertexProcessor.cpp:302-317).  RecordingRasterizer is not a
// swr::Rasterizer: it receives, batch by batch, the host arrays the reference hands to draw*List -- screen-space
// RasterizerVertex records and indices with -1 for dropped primitives -- counts what it sees and forwards the batch
// to an inner swr::Rasterizer (whose draw*List takes such host arrays).  The result must be what the direct path
// draws: VertexProcessorTest's clipped triangle arrives as the 3 fan triangles of its clipped pentagon and
// covers 41 068 pixels; Benchmark.cpp's 40 960 triangles arrive in 40 batches and draw 240 235 639 fragments.
#include <swr/Renderer.h>

#include <cstdio>
#include <vector>

using namespace swr;

struct VertexData {
    float x, y, z;
    float r, g, b;
};

class PixelShader : public PixelShaderBase<PixelShader> {
public:
    static const int AVarCount = 3;
    static const int RenderTargets = 1;
    __device__ static void drawPixel(const PixelData &p) { target<int>(p, 0) = 1; }
};

class VertexShader : public VertexShaderBase<VertexShader> {
public:
    static const int AttribCount = 1;
    static const int AVarCount = 3;
    static const int PVarCount = 0;
    __device__ static void processVertex(VertexShaderInput in, VertexShaderOutput *out)
    {
        const VertexData *data = static_cast<const VertexData *>(in[0]);
        out->x = data->x; out->y = data->y; out->z = data->z; out->w = 1.0f;
        out->avar[0] = data->r; out->avar[1] = data->g; out->avar[2] = data->b;
    }
};

class RecordingRasterizer : public IRasterizer {
public:
    Rasterizer inner;
    mutable long batches = 0, primitives = 0, dropped = 0;

    void drawPointList(const RasterizerVertex *vertices, const int *indices, size_t indexCount) const override
    {
        count(indices, indexCount, 1);
        inner.drawPointList(vertices, indices, indexCount);
    }
    void drawLineList(const RasterizerVertex *vertices, const int *indices, size_t indexCount) const override
    {
        count(indices, indexCount, 2);
        inner.drawLineList(vertices, indices, indexCount);
    }
    void drawTriangleList(const RasterizerVertex *vertices, const int *indices, size_t indexCount) const override
    {
        count(indices, indexCount, 3);
        inner.drawTriangleList(vertices, indices, indexCount);
    }

private:
    void count(const int *indices, size_t indexCount, size_t per) const
    {
        ++batches;
        for (size_t i = 0; i + per <= indexCount; i += per) (indices[i] == -1 ? dropped : primitives)++;
    }
};

// Random.cpp:7-50 (Knuth subtractive generator of System.Random), enough of it for NextDouble().
class Random {
    int seedArray[56];
    int inext = 0, inextp = 21;

public:
    explicit Random(int seed)
    {
        const int MBIG = 2147483647, MSEED = 161803398;
        int mj = MSEED - (seed < 0 ? -seed : seed), mk = 1;
        seedArray[55] = mj;
        for (int i = 1; i < 55; i++) {
            int ii = (21 * i) % 55;
            seedArray[ii] = mk;
            mk = mj - mk;
            if (mk < 0) mk += MBIG;
            mj = seedArray[ii];
        }
        for (int k = 1; k < 5; k++)
            for (int i = 1; i < 56; i++) {
                seedArray[i] -= seedArray[1 + (i + 30) % 55];
                if (seedArray[i] < 0) seedArray[i] += MBIG;
            }
    }
    double NextDouble()
    {
        const int MBIG = 2147483647;
        if (++inext >= 56) inext = 1;
        if (++inextp >= 56) inextp = 1;
        int r = seedArray[inext] - seedArray[inextp];
        if (r == MBIG) r--;
        if (r < 0) r += MBIG;
        seedArray[inext] = r;
        return r * (1.0 / MBIG);
    }
};

static void report(RecordingRasterizer &r, int *buffer, const char *what)
{
    swr_stats st;
    swr_get_stats(r.inner.context(), &st);
    std::vector<int> host(640 * 480);
    swr_memcpy_d2h(r.inner.context(), host.data(), buffer, sizeof(int) * host.size());
    r.inner.finish();
    long covered = 0;
    for (int px : host) covered += px;
    std::printf("%s: batches %ld primitives %ld dropped %ld fragments %llu covered %ld\n", what, r.batches, r.primitives, r.dropped,
                (unsigned long long)st.fragments, covered);
}

int main()
{
    RecordingRasterizer r;
    VertexProcessor v(&r);                                   // not a swr::Rasterizer: the IRasterizer plug point

    r.inner.setScissorRect(0, 0, 640, 480);
    r.inner.setPixelShader<PixelShader>();
    int *buffer = static_cast<int *>(swr_device_alloc(r.inner.context(), sizeof(int) * 640 * 480));
    swr_memset32(r.inner.context(), buffer, 0, 640 * 480);
    r.inner.setRenderTarget(0, buffer, 640 * 4, 640, 480);

    v.setCullMode(CullMode::None);
    v.setVertexShader<VertexShader>();

    // VertexProcessorTest.cpp:87-120
    v.setViewport(100, 100, 640 - 200, 480 - 200);
    VertexData vdata[3] = { { 0.0f, 0.5f, 0.0f, 1, 0, 0 }, { -1.5f, -0.5f, 0.0f, 0, 1, 0 }, { 1.5f, -0.5f, 0.0f, 0, 0, 1 } };
    int idata[3] = { 0, 1, 2 };
    v.setVertexAttribPointer(0, sizeof(VertexData), vdata);
    v.drawElements(DrawMode::Triangle, 3, idata);
    report(r, buffer, "vertex_processor_test");

    // Benchmark.cpp:66-102
    r.batches = r.primitives = r.dropped = 0;
    swr_reset_stats(r.inner.context());
    swr_memset32(r.inner.context(), buffer, 0, 640 * 480);
    v.setViewport(0, 0, 640, 480);
    std::vector<int> indices;
    std::vector<VertexData> vertices;
    Random random(0);
    for (int i = 0; i < 4096 * 10; i++) {
        for (int k = 0; k < 3; ++k) {
            VertexData d;
            d.x = (float)random.NextDouble(); d.y = (float)random.NextDouble(); d.z = (float)random.NextDouble();
            d.r = (float)random.NextDouble(); d.g = (float)random.NextDouble(); d.b = (float)random.NextDouble();
            indices.push_back((int)vertices.size());
            vertices.push_back(d);
        }
    }
    v.setVertexAttribPointer(0, sizeof(VertexData), &vertices[0]);
    v.drawElements(DrawMode::Triangle, indices.size(), &indices[0]);
    report(r, buffer, "benchmark");
    swr_device_free(r.inner.context(), buffer);
    return 0;
}
