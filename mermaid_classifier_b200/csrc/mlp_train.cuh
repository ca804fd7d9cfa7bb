// K9: MLP training step kernels (placeholder).
#pragma once
#include "common.cuh"
