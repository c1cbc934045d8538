/* Minimal fixed-function OpenGL stand-in so the reference's translation units
 * compile on a box with no GL development package.  Every entry point is an
 * empty inline: the headless (--nogfx) path never draws.  Test infrastructure
 * only (oracle build); nothing here is shipped in the product. */
#pragma once
typedef float GLfloat;
typedef unsigned char GLubyte;
typedef unsigned int GLenum;
typedef unsigned int GLbitfield;
typedef int GLint;
typedef int GLsizei;
enum {
    GL_POINTS = 0x0000, GL_LINES = 0x0001, GL_LINE_LOOP = 0x0002, GL_QUADS = 0x0007,
    GL_DEPTH_BUFFER_BIT = 0x0100, GL_COLOR_BUFFER_BIT = 0x4000,
    GL_DEPTH_TEST = 0x0B71, GL_LIGHTING = 0x0B50, GL_COLOR_MATERIAL = 0x0B57,
    GL_LIGHT0 = 0x4000 + 1, GL_POSITION = 0x1203,
    GL_MODELVIEW = 0x1700, GL_PROJECTION = 0x1701
};
static inline void glBegin(GLenum) {}
static inline void glEnd() {}
static inline void glColor3f(GLfloat, GLfloat, GLfloat) {}
static inline void glDisable(GLenum) {}
static inline void glEnable(GLenum) {}
static inline void glLineWidth(GLfloat) {}
static inline void glPointSize(GLfloat) {}
static inline void glVertex3fv(const GLfloat*) {}
static inline void glNormal3fv(const GLfloat*) {}
static inline void glMultMatrixf(const GLfloat*) {}
static inline void glPushMatrix() {}
static inline void glPopMatrix() {}
static inline void glScalef(GLfloat, GLfloat, GLfloat) {}
static inline void glTranslatef(GLfloat, GLfloat, GLfloat) {}
static inline void glViewport(GLint, GLint, GLsizei, GLsizei) {}
static inline void glClearColor(GLfloat, GLfloat, GLfloat, GLfloat) {}
static inline void glClear(GLbitfield) {}
static inline void glMatrixMode(GLenum) {}
static inline void glLoadIdentity() {}
static inline void glLightfv(GLenum, GLenum, const GLfloat*) {}
