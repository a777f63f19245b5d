// examples/benchmark.cu -- the reference's src/examples/Benchmark.cpp, written against the B200 API.
//
// The program structure, the shader classes (CRTP on swr::VertexShaderBase / swr::PixelShaderBase), the
// state setters and the drawElements call are the reference's; what changes is listed in
// drawPixel / processVertex are __device__, the frame buffer is a registered render
// target reached through swr::target<>, and the file is compiled by nvcc and linked with libswr_b200.so.
// Prints "Elapsed: <ms>" like the original, plus the fragment count (240235639 for Span, the
// reference's own number).
#include <swr/Renderer.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <vector>

using namespace swr;

struct VertexData {
    float x, y, z;
    float r, g, b;
};

struct PixelShader : public PixelShaderBase<PixelShader> {
    static const int AVarCount = 3;
    static const int RenderTargets = 1;      // additive: slot 0 is staged per tile

    __device__ static void drawPixel(const PixelData &p)
    {
        target<int>(p, 0) = 1;               // Benchmark.cpp:22-25: buffer[p.x + width * p.y] = 1
    }
};

struct VertexShader : public VertexShaderBase<VertexShader> {
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

// Random.cpp:7-50 (Knuth subtractive generator of System.Random), enough of it for NextDouble().
class Random {
    int seedArray[56];
    int inext = 0, inextp = 21;

public:
    explicit Random(int seed)
    {
        const int MBIG = 2147483647, MSEED = 161803398;
        int mj = MSEED - std::abs(seed), mk = 1;
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

static VertexData CreateVertex(Random &random)
{
    VertexData v;
    v.x = (float)random.NextDouble();
    v.y = (float)random.NextDouble();
    v.z = (float)random.NextDouble();
    v.r = (float)random.NextDouble();
    v.g = (float)random.NextDouble();
    v.b = (float)random.NextDouble();
    return v;
}

int main(int argc, char *argv[])
{
    const int width = 640, height = 480;
    Rasterizer r;
    VertexProcessor v(&r);

    r.setScissorRect(0, 0, width, height);
    v.setViewport(0, 0, width, height);
    v.setCullMode(CullMode::None);
    if (argc > 1 && argv[1][0] == 'b') r.setRasterMode(RasterMode::Block);

    // the frame buffer lives in device memory and is registered as render target 0
    int *buffer = static_cast<int *>(swr_device_alloc(r.context(), sizeof(int) * width * height));
    swr_memset32(r.context(), buffer, 0, (size_t)width * height);
    r.setRenderTarget(0, buffer, width * 4, width, height);

    std::vector<int> indices;
    std::vector<VertexData> vertices;
    Random random(0);
    for (int i = 0; i < 4096 * 10; i++) {
        int offset = (int)vertices.size();
        vertices.push_back(CreateVertex(random));
        vertices.push_back(CreateVertex(random));
        vertices.push_back(CreateVertex(random));
        indices.push_back(offset + 0);
        indices.push_back(offset + 1);
        indices.push_back(offset + 2);
    }

    r.setPixelShader<PixelShader>();
    v.setVertexShader<VertexShader>();
    v.setVertexAttribPointer(0, sizeof(VertexData), &vertices[0]);          // Benchmark.cpp:99, as is: no extent

    v.drawElements(DrawMode::Triangle, indices.size(), &indices[0]);     // warm-up (allocates scratch)
    swr_reset_stats(r.context());

    auto start = std::chrono::steady_clock::now();
    v.drawElements(DrawMode::Triangle, indices.size(), &indices[0]);
    auto end = std::chrono::steady_clock::now();
    std::cout << "Elapsed: " << std::chrono::duration_cast<std::chrono::microseconds>(end - start).count() / 1000.0 << std::endl;

    swr_stats st;
    swr_get_stats(r.context(), &st);
    std::vector<int> host(width * height);
    swr_memcpy_d2h(r.context(), host.data(), buffer, sizeof(int) * host.size());
    r.finish();
    long covered = 0;
    for (int px : host) covered += px;
    std::printf("fragments %llu covered %ld kernels %llu\n", (unsigned long long)st.fragments, covered,
                (unsigned long long)st.kernel_launches);
    swr_device_free(r.context(), buffer);
    return 0;
}
