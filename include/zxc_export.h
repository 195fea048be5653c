/*
 * zxc_export.h -- symbol visibility for the B200 drop-in libzxc.
 *
 * Replaces: reference include/zxc_export.h:60-80 (ZXC_EXPORT / ZXC_NO_EXPORT /
 * ZXC_DEPRECATED macros).  The library is built with -fvisibility=hidden and
 * only ZXC_EXPORT symbols are visible, as in the reference (CMakeLists.txt:84-92).
 */
#ifndef ZXC_EXPORT_H
#define ZXC_EXPORT_H

#if defined(ZXC_STATIC_DEFINE)
#  define ZXC_EXPORT
#  define ZXC_NO_EXPORT
#elif defined(_WIN32)
#  if defined(zxc_lib_EXPORTS)
#    define ZXC_EXPORT __declspec(dllexport)
#  elif defined(ZXC_DLL_IMPORT)
#    define ZXC_EXPORT __declspec(dllimport)
#  else
#    define ZXC_EXPORT
#  endif
#  define ZXC_NO_EXPORT
#else
#  define ZXC_EXPORT __attribute__((visibility("default")))
#  define ZXC_NO_EXPORT __attribute__((visibility("hidden")))
#endif

#if defined(_WIN32)
#  define ZXC_DEPRECATED __declspec(deprecated)
#else
#  define ZXC_DEPRECATED __attribute__((__deprecated__))
#endif
#define ZXC_DEPRECATED_EXPORT ZXC_EXPORT ZXC_DEPRECATED
#define ZXC_DEPRECATED_NO_EXPORT ZXC_NO_EXPORT ZXC_DEPRECATED

#endif /* ZXC_EXPORT_H */
