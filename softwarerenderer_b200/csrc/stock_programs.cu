// stock_programs.cu -- instantiates the geometry / tile kernels for the stock shader pack.
//
// Built once per unit (-DSWR_STOCK_UNIT=n, see Makefile) so the heavy tile-kernel instantiations
// compile in parallel: unit 0 = the vertex shaders, units 1..7 = one pixel shader each.  Every
// unit is its own translation unit and therefore has its own uniform block (swr/Uniforms.h).
#include "stock_shaders.cuh"

#ifndef SWR_STOCK_UNIT
#error "build with -DSWR_STOCK_UNIT=n"
#endif

using namespace swr::detail;

#if SWR_STOCK_UNIT == 0
extern "C" SWR_API const swr_vertex_shader *swr_stock_vertex_shader(int vs_kind)
{
    switch (vs_kind) {
    case SWR_VS_POS_COLOR: return vertexShaderBinding<stock::VSPosColor>("pos_color");
    case SWR_VS_MVP_COLOR: return vertexShaderBinding<stock::VSMvpColor>("mvp_color");
    case SWR_VS_MVP_NORMAL_UV: return vertexShaderBinding<stock::VSMvpNormalUv>("mvp_normal_uv");
    default: return nullptr;
    }
}

extern "C" const swr_pixel_shader *swr_stock_ps_flat(void);
extern "C" const swr_pixel_shader *swr_stock_ps_count_id(void);
extern "C" const swr_pixel_shader *swr_stock_ps_gouraud(void);
extern "C" const swr_pixel_shader *swr_stock_ps_gouraud_depth(void);
extern "C" const swr_pixel_shader *swr_stock_ps_vary_dump(void);
extern "C" const swr_pixel_shader *swr_stock_ps_textured(void);
extern "C" const swr_pixel_shader *swr_stock_ps_textured_aniso(void);

extern "C" SWR_API const swr_pixel_shader *swr_stock_pixel_shader(int ps_kind)
{
    switch (ps_kind) {
    case SWR_PS_FLAT: return swr_stock_ps_flat();
    case SWR_PS_COUNT_ID: return swr_stock_ps_count_id();
    case SWR_PS_GOURAUD: return swr_stock_ps_gouraud();
    case SWR_PS_GOURAUD_DEPTH: return swr_stock_ps_gouraud_depth();
    case SWR_PS_VARY_DUMP: return swr_stock_ps_vary_dump();
    case SWR_PS_TEXTURED: return swr_stock_ps_textured();
    case SWR_PS_TEXTURED_ANISO: return swr_stock_ps_textured_aniso();
    default: return nullptr;
    }
}
#elif SWR_STOCK_UNIT == 1
extern "C" const swr_pixel_shader *swr_stock_ps_flat(void) { return pixelShaderBinding<stock::PSFlat>("flat"); }
#elif SWR_STOCK_UNIT == 2
extern "C" const swr_pixel_shader *swr_stock_ps_count_id(void) { return pixelShaderBinding<stock::PSCountId>("count_id"); }
#elif SWR_STOCK_UNIT == 3
extern "C" const swr_pixel_shader *swr_stock_ps_gouraud(void) { return pixelShaderBinding<stock::PSGouraud>("gouraud"); }
#elif SWR_STOCK_UNIT == 4
extern "C" const swr_pixel_shader *swr_stock_ps_gouraud_depth(void) { return pixelShaderBinding<stock::PSGouraudDepth>("gouraud_depth"); }
#elif SWR_STOCK_UNIT == 5
extern "C" const swr_pixel_shader *swr_stock_ps_vary_dump(void) { return pixelShaderBinding<stock::PSVaryDump>("vary_dump"); }
#elif SWR_STOCK_UNIT == 6
extern "C" const swr_pixel_shader *swr_stock_ps_textured(void) { return pixelShaderBinding<stock::PSTextured>("textured"); }
#elif SWR_STOCK_UNIT == 7
extern "C" const swr_pixel_shader *swr_stock_ps_textured_aniso(void) { return pixelShaderBinding<stock::PSTexturedAniso>("textured_aniso"); }
#endif
