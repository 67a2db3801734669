/* Hand-written build configuration for compiling the UNMODIFIED reference
 * runtime sources (where they lie under /root/reference) into oracle/_ref/.
 * Plays the role of the file CMake would generate from cmake/config.h.in
 * (reference: cmake/config.h.in:1-110).  TEST INFRASTRUCTURE ONLY. */
#ifndef CONFIG_H_
#define CONFIG_H_
#define ARCH_64_BIT
#define ARCH_X86_64
#define HAVE_CC_BUILTIN_ASSUME_ALIGNED
#define HAVE_CXX_BUILTIN_ASSUME_ALIGNED
#define HAVE_C_X86INTRIN_H
#define HAVE_CXX_X86INTRIN_H
#define HAVE_POSIX_MEMALIGN
#define HAVE_UNISTD_H
#define HAVE__BUILTIN_CONSTANT_P
#define HS_OPTIMIZE
#define RELEASE_BUILD
#endif
