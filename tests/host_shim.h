// Minimal stand-ins so the CUDA headers compile as plain C++ for the CPU-side arithmetic tests.
#pragma once
struct uint4 { unsigned int x, y, z, w; };
