/* GL/gl.h stand-in for compiling the reference's parsers in the tests: the scalar typedefs only. */
#ifndef PBR_REF_GL_H
#define PBR_REF_GL_H
typedef float GLfloat;
typedef int GLint;
typedef unsigned int GLuint;
typedef unsigned char GLubyte;
typedef unsigned int GLenum;
#endif
