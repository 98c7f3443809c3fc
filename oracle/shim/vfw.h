/* Minimal <vfw.h> stand-in (see windows.h in this directory). */
#ifndef ORACLE_SHIM_VFW_H
#define ORACLE_SHIM_VFW_H
#include <windows.h>
#define mmioFOURCC(a,b,c,d) ((DWORD)(BYTE)(a) | ((DWORD)(BYTE)(b) << 8) | ((DWORD)(BYTE)(c) << 16) | ((DWORD)(BYTE)(d) << 24))
typedef struct { DWORD dwFlags; BITMAPINFOHEADER *lpbiOutput; LPVOID lpOutput; BITMAPINFOHEADER *lpbiInput;
                 LPVOID lpInput; DWORD *lpckid; DWORD *lpdwFlags; LONG lFrameNum; DWORD dwFrameSize;
                 DWORD dwQuality; BITMAPINFOHEADER *lpbiPrev; LPVOID lpPrev; } ICCOMPRESS;
typedef struct { DWORD dwFlags; BITMAPINFOHEADER *lpbiOutput; LPARAM lOutput; BITMAPINFOHEADER *lpbiInput;
                 LPARAM lInput; LONG lStartFrame; LONG lFrameCount; LONG lQuality; LONG lDataRate;
                 LONG lKeyRate; DWORD dwRate; DWORD dwScale; DWORD dwOverheadPerFrame; DWORD dwReserved2;
                 void *GetData; void *PutData; } ICCOMPRESSFRAMES;
typedef struct { DWORD dwFlags; BITMAPINFOHEADER *lpbiInput; LPVOID lpInput; BITMAPINFOHEADER *lpbiOutput;
                 LPVOID lpOutput; DWORD ckid; } ICDECOMPRESS;
#endif
