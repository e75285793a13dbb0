// Stand-in for <GL/gl.h>: the reference's src/render_buffer.cu includes <cuda_gl_interop.h> even with the GUI disabled,
// and that CUDA header only needs these three typedefs.  Written for oracle/Makefile.ref; not part of the product.
#pragma once
typedef unsigned int GLenum;
typedef unsigned int GLuint;
typedef int GLint;
