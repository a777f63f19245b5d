// swr/Renderer.h -- umbrella include, as the reference's src/renderer/Renderer.h:27-28.
#pragma once

#include "Rasterizer.h"
#include "VertexProcessor.h"
