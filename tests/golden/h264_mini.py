"""A minimal H.264 Annex-B WRITER (ITU-T H.264 7.3: SPS, PPS, slice header, CAVLC slice data) -- test
infrastructure for tests/golden/make_h264_pins.py.  It writes only what is needed to make a standard decoder
evaluate the NORMATIVE prediction processes on pixels we choose, with no residual and no deblocking:

  * I_PCM macroblocks (7.3.5: mb_type 25, raw samples) carry the chosen pixels;
  * P_L0_16x16 macroblocks with one motion vector for the whole picture and coded_block_pattern 0 decode to the
    motion-compensated prediction alone = the fractional sample interpolation of 8.4.2.2.1 on the I_PCM picture
    (what [x264] hpel_filter + get_ref must reproduce: the encoder's prediction has to be the decoder's);
  * I_NxN macroblocks with transform_size_8x8_flag and coded_block_pattern 0 decode to Intra_8x8 prediction
    (8.3.2, incl. the reference sample filtering) and Intra chroma prediction (8.3.4) alone.

Nothing here comes from the reference tree (the reference holds no bitstream writer); it follows the standard's
syntax tables."""


class Bits:
    """MSB-first bit writer; whole bytes go in at byte speed once aligned (I_PCM samples)."""

    def __init__(self):
        self.buf = bytearray()
        self.acc = 0
        self.n = 0

    def u(self, n, v):
        self.acc = (self.acc << n) | (v & ((1 << n) - 1))
        self.n += n
        while self.n >= 8:
            self.n -= 8
            self.buf.append((self.acc >> self.n) & 0xFF)
        self.acc &= (1 << self.n) - 1

    def ue(self, v):
        v += 1
        n = v.bit_length()
        self.u(n - 1, 0)
        self.u(n, v)

    def se(self, v):
        self.ue(2 * v - 1 if v > 0 else -2 * v)

    def aligned(self):
        return self.n == 0

    def align_zero(self):
        if self.n:
            self.u(8 - self.n, 0)

    def raw(self, data: bytes):
        assert self.aligned()
        self.buf += data

    def trailing(self):                     # rbsp_trailing_bits
        self.u(1, 1)
        self.align_zero()

    def bytes(self):
        assert self.aligned()
        return bytes(self.buf)


def nal(ref_idc, unit_type, rbsp: bytes) -> bytes:
    """Start code + header + payload with emulation prevention (7.4.1): 03 after two zero bytes when the next
    byte is 00..03 (the zero run starts again at that byte)."""
    import re
    return b"\x00\x00\x00\x01" + bytes([(ref_idc << 5) | unit_type]) + re.sub(b"\x00\x00(?=[\x00-\x03])", b"\x00\x00\x03", rbsp)


def sps(mb_w, mb_h, num_ref_frames=1, reorder=0) -> bytes:
    b = Bits()
    b.u(8, 100)                  # profile_idc: High (transform 8x8 for the Intra_8x8 pictures)
    b.u(8, 0)                    # constraint flags + reserved
    b.u(8, 40)                   # level_idc
    b.ue(0)                      # seq_parameter_set_id
    b.ue(1)                      # chroma_format_idc 4:2:0
    b.ue(0); b.ue(0)             # bit_depth_luma/chroma_minus8
    b.u(1, 0)                    # qpprime_y_zero_transform_bypass_flag
    b.u(1, 0)                    # seq_scaling_matrix_present_flag
    b.ue(0)                      # log2_max_frame_num_minus4 -> 4 bits
    b.ue(0)                      # pic_order_cnt_type 0
    b.ue(4)                      # log2_max_pic_order_cnt_lsb_minus4 -> 8 bits
    b.ue(num_ref_frames)         # max_num_ref_frames
    b.u(1, 0)                    # gaps_in_frame_num_value_allowed_flag
    b.ue(mb_w - 1); b.ue(mb_h - 1)
    b.u(1, 1)                    # frame_mbs_only_flag
    b.u(1, 1)                    # direct_8x8_inference_flag
    b.u(1, 0)                    # frame_cropping_flag
    if not reorder:
        b.u(1, 0)                # vui_parameters_present_flag
    else:                        # VUI with nothing but the bitstream restriction: tells the decoder its reorder depth
        b.u(1, 1)
        for _ in range(8):       # aspect_ratio, overscan, video_signal_type, chroma_loc, timing, nal_hrd, vcl_hrd,
            b.u(1, 0)            # pic_struct: all absent
        b.u(1, 1)                # bitstream_restriction_flag
        b.u(1, 1)                # motion_vectors_over_pic_boundaries_flag
        b.ue(0); b.ue(0)         # max_bytes_per_pic_denom, max_bits_per_mb_denom
        b.ue(16); b.ue(16)       # log2_max_mv_length_horizontal / vertical
        b.ue(reorder)            # max_num_reorder_frames
        b.ue(num_ref_frames)     # max_dec_frame_buffering
    b.trailing()
    return nal(3, 7, b.bytes())


def pps(pps_id=0, weighted_pred=0, weighted_bipred_idc=0) -> bytes:
    b = Bits()
    b.ue(pps_id); b.ue(0)        # pic_parameter_set_id, seq_parameter_set_id
    b.u(1, 0)                    # entropy_coding_mode_flag: CAVLC
    b.u(1, 0)                    # bottom_field_pic_order_in_frame_present_flag
    b.ue(0)                      # num_slice_groups_minus1
    b.ue(0); b.ue(0)             # num_ref_idx_l0/l1_default_active_minus1
    b.u(1, weighted_pred); b.u(2, weighted_bipred_idc)     # weighted_pred_flag, weighted_bipred_idc (2 = implicit)
    b.se(0); b.se(0); b.se(0)    # pic_init_qp_minus26, pic_init_qs_minus26, chroma_qp_index_offset
    b.u(1, 1)                    # deblocking_filter_control_present_flag (slices switch the filter off)
    b.u(1, 0)                    # constrained_intra_pred_flag
    b.u(1, 0)                    # redundant_pic_cnt_present_flag
    b.u(1, 1)                    # transform_8x8_mode_flag
    b.u(1, 0)                    # pic_scaling_matrix_present_flag
    b.se(0)                      # second_chroma_qp_index_offset
    b.trailing()
    return nal(3, 8, b.bytes())


def _slice_header(b: Bits, idr: bool, slice_type: int, frame_num: int, poc_lsb: int, ref: bool, pps_id=0, weight=None):
    b.ue(0)                      # first_mb_in_slice
    b.ue(slice_type)             # 0 = P, 1 = B, 2 = I
    b.ue(pps_id)                 # pic_parameter_set_id
    b.u(4, frame_num)
    if idr:
        b.ue(0)                  # idr_pic_id
    b.u(8, poc_lsb)
    if slice_type == 1:
        b.u(1, 1)                # direct_spatial_mv_pred_flag
        b.u(1, 0)                # num_ref_idx_active_override_flag
        b.u(1, 0); b.u(1, 0)     # ref_pic_list_modification_flag_l0 / _l1
    if slice_type == 0:
        b.u(1, 0)                # num_ref_idx_active_override_flag
        b.u(1, 0)                # ref_pic_list_modification_flag_l0
        if weight is not None:   # pred_weight_table (7.3.3.2), the PPS has weighted_pred_flag
            denom, scale, offset = weight
            b.ue(denom)          # luma_log2_weight_denom
            b.ue(0)              # chroma_log2_weight_denom
            b.u(1, 1); b.se(scale); b.se(offset)     # luma_weight_l0_flag, luma_weight_l0, luma_offset_l0
            b.u(1, 0)            # chroma_weight_l0_flag
    if ref:
        if idr:
            b.u(1, 0); b.u(1, 0)     # no_output_of_prior_pics_flag, long_term_reference_flag
        else:
            b.u(1, 0)                # adaptive_ref_pic_marking_mode_flag
    b.se(0)                      # slice_qp_delta
    b.ue(1)                      # disable_deblocking_filter_idc = 1: off


def _pcm_mb(b: Bits, mb_type_code: int, y, u, v, mbx, mby):
    b.ue(mb_type_code)           # I_PCM: 25 in I slices, 5 + 25 in P slices
    b.align_zero()               # pcm_alignment_zero_bit
    b.raw(y[16 * mby:16 * mby + 16, 16 * mbx:16 * mbx + 16].tobytes())
    b.raw(u[8 * mby:8 * mby + 8, 8 * mbx:8 * mbx + 8].tobytes())
    b.raw(v[8 * mby:8 * mby + 8, 8 * mbx:8 * mbx + 8].tobytes())


def idr_pcm_picture(y, u, v, intra_tests=None) -> bytes:
    """An IDR picture of I_PCM macroblocks.  intra_tests: {(mbx, mby): (luma Intra_8x8 mode 0..8, chroma mode 0..3)}
    replaces those macroblocks by I_NxN / transform 8x8 with no residual (their neighbours must stay I_PCM)."""
    mb_h, mb_w = y.shape[0] // 16, y.shape[1] // 16
    intra_tests = intra_tests or {}
    b = Bits()
    _slice_header(b, True, 2, 0, 0, True)
    for mby in range(mb_h):
        for mbx in range(mb_w):
            if (mbx, mby) not in intra_tests:
                _pcm_mb(b, 25, y, u, v, mbx, mby)
                continue
            m, c = intra_tests[(mbx, mby)]
            for nb in ((mbx - 1, mby), (mbx, mby - 1), (mbx - 1, mby - 1), (mbx + 1, mby - 1)):
                assert nb not in intra_tests and nb[0] >= 0 and nb[1] >= 0
            b.ue(0)              # mb_type I_NxN
            b.u(1, 1)            # transform_size_8x8_flag
            # predIntra8x8PredMode (8.3.2.1) = min(mode of the left block, mode of the block above); an I_PCM
            # neighbour counts as DC (2).  Block 0 sees two I_PCM neighbours, blocks 1 and 2 one, block 3 sees m twice.
            for blk in range(4):
                pred = 2 if blk == 0 else min(m, 2) if blk < 3 else m
                if pred == m:
                    b.u(1, 1)    # prev_intra8x8_pred_mode_flag
                else:
                    b.u(1, 0)
                    b.u(3, m if m < pred else m - 1)     # rem_intra8x8_pred_mode
            b.ue(c)              # intra_chroma_pred_mode: 0 DC, 1 horizontal, 2 vertical, 3 plane
            b.ue(3)              # coded_block_pattern me(v): codeNum 3 = Intra cbp 0 (table 9-4)
    b.trailing()
    return nal(3, 5, b.bytes())


def p_picture_uniform_mv(mb_w, mb_h, mvx, mvy, frame_num, poc_lsb, weight=None) -> bytes:
    """A non-reference P picture: every macroblock P_L0_16x16 with the quarter-sample vector (mvx, mvy), no
    residual.  weight = (log2 denominator, scale, offset): explicit weighted prediction of luma (needs pps(1, 1),
    selected here by pps id 1).  With one vector everywhere the median predictor (8.4.1.3) equals it for every macroblock but the
    first, whose neighbours are all unavailable (predictor 0)."""
    b = Bits()
    _slice_header(b, False, 0, frame_num, poc_lsb, False, pps_id=1 if weight is not None else 0, weight=weight)
    for i in range(mb_w * mb_h):
        b.ue(0)                  # mb_skip_run
        b.ue(0)                  # mb_type P_L0_16x16
        b.se(mvx if i == 0 else 0)
        b.se(mvy if i == 0 else 0)
        b.ue(0)                  # coded_block_pattern me(v): codeNum 0 = Inter cbp 0
    b.trailing()
    return nal(0, 1, b.bytes())


def p_pcm_reference_picture(y, u, v, frame_num, poc_lsb) -> bytes:
    """A P picture made of I_PCM macroblocks, kept as a reference: a second reference with pixels we choose."""
    mb_h, mb_w = y.shape[0] // 16, y.shape[1] // 16
    b = Bits()
    _slice_header(b, False, 0, frame_num, poc_lsb, True)
    for mby in range(mb_h):
        for mbx in range(mb_w):
            b.ue(0)              # mb_skip_run
            _pcm_mb(b, 30, y, u, v, mbx, mby)
    b.trailing()
    return nal(2, 1, b.bytes())


def b_picture_uniform_mvs(mb_w, mb_h, mv0, mv1, frame_num, poc_lsb, pps_id) -> bytes:
    """A non-reference B picture: every macroblock B_Bi_16x16 (list 0 = the nearest earlier reference, list 1 = the
    nearest later one) with one vector per list, no residual: the decoded picture is the (pps: default or implicitly
    weighted) average of the two interpolated predictions."""
    b = Bits()
    _slice_header(b, False, 1, frame_num, poc_lsb, False, pps_id=pps_id)
    for i in range(mb_w * mb_h):
        b.ue(0)                  # mb_skip_run
        b.ue(3)                  # mb_type B_Bi_16x16
        for mv in (mv0, mv1):    # mvd_l0 then mvd_l1; per list the median predictor is the uniform vector
            b.se(mv[0] if i == 0 else 0)
            b.se(mv[1] if i == 0 else 0)
        b.ue(0)                  # coded_block_pattern: Inter 0
    b.trailing()
    return nal(0, 1, b.bytes())
