/* empty: no HAVE_FFMPEG etc. (the reference copies config.h from an x264 build dir, Makefile:57-73) */
