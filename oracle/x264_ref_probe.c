/*
 * x264_ref_probe.c -- differential probe against a REAL libx264, if one is ever provided (SURVEY.md 0.1:
 * baseline/_ref/ is reserved for "the reference install").  TEST INFRASTRUCTURE ONLY.  Built by
 * `make -C oracle x264probe` only when an x264.h and a libx264 are found under baseline/_ref or oracle/_ref;
 * there is none in this image, so this file has never been compiled against the real header (it uses the public
 * API only, the same calls the reference makes: codec.c:1463 x264_param_default_preset, :1349 x264_param_parse,
 * :1584 x264_param_apply_profile, :1623 x264_encoder_open, :1693 x264_encoder_encode, :1848 delayed frames,
 * :1857 x264_encoder_close).
 *
 * usage: x264_ref_probe <width> <height> <frames> <preset> [key=value ...] < tight I420 frames
 * prints one line per coded frame: "<display index> <X264_TYPE_* of pic_out.i_type> <b_keyframe>"
 * tests/test_x264_differential.py (skipped while the binary is absent) compares those frame types with the CPU
 * checker's and the device path's decisions on the same planes: north_star's ">= 99.9 % of frames" criterion.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <x264.h>

int main(int argc, char **argv)
{
    if (argc < 5) { fprintf(stderr, "usage: %s w h frames preset [key=value ...] < i420\n", argv[0]); return 2; }
    const int w = atoi(argv[1]), h = atoi(argv[2]), n = atoi(argv[3]);
    x264_param_t param;
    if (x264_param_default_preset(&param, argv[4], NULL) < 0) return 1;
    param.i_width = w; param.i_height = h; param.i_csp = X264_CSP_I420;
    param.i_fps_num = 25; param.i_fps_den = 1; param.b_vfr_input = 0;        /* codec.c:1567-1569: CFR timebase */
    param.i_threads = 1; param.i_lookahead_threads = 1;                       /* pin the band split */
    param.rc.i_rc_method = X264_RC_CRF; param.rc.f_rf_constant = 23;          /* the wrapper's default (config.c:109-111) */
    param.i_log_level = X264_LOG_ERROR;
    for (int i = 5; i < argc; i++) {
        char *eq = strchr(argv[i], '=');
        if (!eq) continue;
        *eq = 0;
        if (x264_param_parse(&param, argv[i], eq + 1) < 0) { fprintf(stderr, "bad option %s\n", argv[i]); return 1; }
    }
    x264_t *enc = x264_encoder_open(&param);
    if (!enc) return 1;
    x264_picture_t pic, out;
    if (x264_picture_alloc(&pic, X264_CSP_I420, w, h) < 0) return 1;
    x264_nal_t *nal; int i_nal;
    for (int f = 0; f < n; f++) {
        for (int p = 0; p < 3; p++) {
            const int pw = p ? w / 2 : w, ph = p ? h / 2 : h;
            for (int y = 0; y < ph; y++)
                if (fread(pic.img.plane[p] + (size_t)y * pic.img.i_stride[p], 1, pw, stdin) != (size_t)pw) { fprintf(stderr, "short read\n"); return 1; }
        }
        pic.i_pts = f; pic.i_type = X264_TYPE_AUTO;
        if (x264_encoder_encode(enc, &nal, &i_nal, &pic, &out) > 0) printf("%d %d %d\n", (int)out.i_pts, out.i_type, out.b_keyframe);
    }
    while (x264_encoder_delayed_frames(enc) > 0)
        if (x264_encoder_encode(enc, &nal, &i_nal, NULL, &out) > 0) printf("%d %d %d\n", (int)out.i_pts, out.i_type, out.b_keyframe);
    x264_encoder_close(enc);
    x264_picture_clean(&pic);
    return 0;
}
