/* Stand-in for the CMake-generated hs_version.h (reference:
 * src/hs_version.h.in:32-40; version from CMakeLists.txt:4-7 = 5.4.2). */
#ifndef HS_VERSION_H_ORACLE
#define HS_VERSION_H_ORACLE
#define HS_VERSION_STRING "5.4.2 oracle-ref"
#define HS_VERSION_32BIT ((5 << 24) | (4 << 16) | (2 << 8) | 0)
#endif
