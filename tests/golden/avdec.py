"""ctypes access to the H.264 DECODER of the FFmpeg build that ships inside the opencv wheel of this image
(libavcodec 62 / libavutil 60; no headers, so the few struct offsets used are spelled out).  Test infrastructure
for tests/golden/make_h264_pins.py and the optional live check in tests/test_h264_pins.py."""
import ctypes as C
import glob
import os

import numpy as np

EAGAIN = -11


def _libs():
    import cv2                                          # resolves the wheel's private library directory
    d = os.path.join(os.path.dirname(os.path.dirname(cv2.__file__)), "opencv_python_headless.libs")
    avc = C.CDLL(glob.glob(os.path.join(d, "libavcodec-*.so*"))[0])
    avu = C.CDLL(glob.glob(os.path.join(d, "libavutil-*.so*"))[0])
    return avc, avu


def available():
    try:
        _libs()
        return True
    except Exception:
        return False


def decode_h264(access_units):
    """access_units: list of bytes (Annex B, one picture each).  Returns the decoded pictures in output order as
    (Y, U, V) uint8 arrays."""
    avc, avu = _libs()
    P = C.c_void_p
    avc.avcodec_find_decoder_by_name.restype = P
    avc.avcodec_find_decoder_by_name.argtypes = [C.c_char_p]
    avc.avcodec_alloc_context3.restype = P
    avc.avcodec_alloc_context3.argtypes = [P]
    avc.avcodec_open2.argtypes = [P, P, P]
    avc.av_packet_alloc.restype = P
    avc.av_new_packet.argtypes = [P, C.c_int]
    avc.av_packet_unref.argtypes = [P]
    avc.avcodec_send_packet.argtypes = [P, P]
    avc.avcodec_receive_frame.argtypes = [P, P]
    avc.avcodec_free_context.argtypes = [C.POINTER(P)]
    avc.av_packet_free.argtypes = [C.POINTER(P)]
    avu.av_frame_alloc.restype = P
    avu.av_frame_unref.argtypes = [P]
    avu.av_frame_free.argtypes = [C.POINTER(P)]

    codec = avc.avcodec_find_decoder_by_name(b"h264")
    if not codec:
        raise RuntimeError("no h264 decoder in this libavcodec")
    ctx = P(avc.avcodec_alloc_context3(codec))
    if avc.avcodec_open2(ctx, codec, None) < 0:
        raise RuntimeError("avcodec_open2 failed")
    pkt = P(avc.av_packet_alloc())
    frame = P(avu.av_frame_alloc())
    out = []

    def drain():
        while True:
            rc = avc.avcodec_receive_frame(ctx, frame)
            if rc < 0:
                return rc
            # AVFrame: uint8_t *data[8] @0, int linesize[8] @64, uint8_t **extended_data @96, int width @104, height @108
            data = (C.c_void_p * 8).from_address(frame.value)
            ls = (C.c_int * 8).from_address(frame.value + 64)
            w = C.c_int.from_address(frame.value + 104).value
            h = C.c_int.from_address(frame.value + 108).value
            planes = []
            for i, (pw, ph) in enumerate(((w, h), (w // 2, h // 2), (w // 2, h // 2))):
                buf = np.ctypeslib.as_array(C.cast(data[i], C.POINTER(C.c_uint8)), shape=(ph * ls[i],))
                planes.append(buf.reshape(ph, ls[i])[:, :pw].copy())
            out.append(tuple(planes))
            avu.av_frame_unref(frame)

    for au in access_units:
        if avc.av_new_packet(pkt, len(au)) < 0:
            raise RuntimeError("av_new_packet failed")
        # AVPacket: AVBufferRef *buf @0, int64 pts @8, int64 dts @16, uint8_t *data @24, int size @32
        C.memmove(C.c_void_p.from_address(pkt.value + 24).value, au, len(au))
        rc = avc.avcodec_send_packet(ctx, pkt)
        avc.av_packet_unref(pkt)
        if rc < 0:
            raise RuntimeError(f"avcodec_send_packet failed: {rc}")
        drain()
    avc.avcodec_send_packet(ctx, None)                  # flush
    drain()
    avu.av_frame_free(C.byref(frame))
    avc.av_packet_free(C.byref(pkt))
    avc.avcodec_free_context(C.byref(ctx))
    return out
