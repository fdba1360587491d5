#ifndef GT_CONFIG_H
#define GT_CONFIG_H
#define GT_CC "/opt/gcc/bin/gcc"
#define GT_CFLAGS "-O3"
#define GT_CPPFLAGS ""
#define GT_VERSION "1.5.11"
#define GT_MAJOR_VERSION 1
#define GT_MINOR_VERSION 5
#define GT_MICRO_VERSION 11
#endif
