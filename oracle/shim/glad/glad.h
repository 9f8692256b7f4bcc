// GL type / enum stubs (OpenGL is replaced by CUDA in this repository; nothing here executes).
#pragma once
#include <cstdint>
typedef unsigned int GLuint;
typedef unsigned int GLenum;
typedef int GLint;
#define GL_RGBA8 0x8058
#define GL_RGBA16F 0x881A
#define GL_R32F 0x822E
#define GL_RG16F 0x822F
#define GL_R8UI 0x8232
#define GL_DYNAMIC_STORAGE_BIT 0x0100
#define GL_MAP_WRITE_BIT 0x0002
