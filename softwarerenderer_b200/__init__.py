"""softwarerenderer_b200 -- a B200-native (sm_100a) implementation of trenki2/SoftwareRenderer's
draw path behind the reference's own API.

  softwarerenderer_b200.api     Rasterizer / VertexProcessor mirror over the C ABI (needs the built
                                libswr_b200.so; importing it without the CUDA extension raises)
  softwarerenderer_b200.scenes  synthetic meshes, cameras and the BASELINE.json configs (numpy only)
  softwarerenderer_b200.dist    sort-first tile partition + NCCL framebuffer composite
  csrc/, ../include/            the CUDA kernels, the C ABI and the C++ CRTP shader API
"""
__version__ = "0.1.0"
