/* Minimal <windows.h> stand-in so that the reference's csp.c compiles UNMODIFIED
 * on Linux for test purposes (oracle/_ref build only).  Test infrastructure:
 * only the types the reference headers mention; nothing here is functional. */
#ifndef ORACLE_SHIM_WINDOWS_H
#define ORACLE_SHIM_WINDOWS_H
#include <stdint.h>
#include <stddef.h>
#include <stdarg.h>
#include <wchar.h>
typedef uint32_t DWORD;
typedef int32_t  LONG;
typedef uint16_t WORD;
typedef uint8_t  BYTE;
typedef int      BOOL;
typedef unsigned UINT;
typedef intptr_t LRESULT, LPARAM, INT_PTR, LONG_PTR;
typedef uintptr_t WPARAM, DWORD_PTR, UINT_PTR;
typedef void *HWND, *HINSTANCE, *HDRVR, *LPVOID, *HANDLE;
typedef char *LPTSTR, *LPSTR;
typedef struct { void *opaque[6]; } CRITICAL_SECTION;
#define CALLBACK
#define WINAPI
#define MAX_PATH 260
#define BI_RGB 0
typedef struct {
    DWORD biSize; LONG biWidth; LONG biHeight; WORD biPlanes; WORD biBitCount;
    DWORD biCompression; DWORD biSizeImage; LONG biXPelsPerMeter; LONG biYPelsPerMeter;
    DWORD biClrUsed; DWORD biClrImportant;
} BITMAPINFOHEADER;
typedef struct { BYTE b, g, r, x; } RGBQUAD;
typedef struct { BITMAPINFOHEADER bmiHeader; RGBQUAD bmiColors[1]; } BITMAPINFO;
#endif
