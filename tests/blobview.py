"""Test helper: decode a packed blob (csrc/blob.h) with numpy."""
import numpy as np

HEADER = np.dtype([("magic", "<u4"), ("count", "<i4"), ("n_cap", "<i4"), ("e_cap", "<i4"), ("total_bytes", "<u8"),
                   ("off_desc", "<u8"), ("off_x", "<u8"), ("off_num", "<u8"), ("off_cur", "<u8"),
                   ("off_rowptr", "<u8"), ("off_adj", "<u8"), ("off_cand_uv", "<u8"), ("off_cand_idx", "<u8"),
                   ("sum_n", "<u8"), ("sum_e", "<u8"), ("sum_k", "<u8"), ("off_order", "<u8"), ("reserved", "<u8", (1,))])
DESC = np.dtype([("n", "<i4"), ("e", "<i4"), ("stage", "<i4"), ("k", "<i4"), ("x_row", "<i4"), ("rp_off", "<i4"),
                 ("adj_off", "<i4"), ("cand_off", "<i4"), ("cost", "<i4"), ("ord_off", "<i4"), ("ord_rounds", "<i4"), ("pad", "<i4", (5,))])
assert HEADER.itemsize == 128 and DESC.itemsize == 64


def decode(buf: np.ndarray):
    buf = np.asarray(buf, dtype=np.uint8)
    h = buf[:128].view(HEADER)[0]
    assert h["magic"] == 0x55504232
    cnt = int(h["count"])
    desc = buf[int(h["off_desc"]):int(h["off_desc"]) + 64 * cnt].view(DESC)
    x = buf[int(h["off_x"]):int(h["off_x"]) + int(h["sum_n"]) * 96].view(np.float32).reshape(-1, 24)
    num = buf[int(h["off_num"]):int(h["off_num"]) + cnt * 208].view(np.float32).reshape(cnt, 52)
    cur = buf[int(h["off_cur"]):int(h["off_cur"]) + cnt * 96].view(np.float32).reshape(cnt, 24)
    rowptr = buf[int(h["off_rowptr"]):int(h["off_order"])].view(np.uint16)
    order = buf[int(h["off_order"]):int(h["off_adj"])].view(np.uint16)
    adj = buf[int(h["off_adj"]):int(h["off_cand_uv"])].view(np.uint32)
    cuv = buf[int(h["off_cand_uv"]):int(h["off_cand_idx"])].view(np.uint32)
    cidx = buf[int(h["off_cand_idx"]):int(h["total_bytes"])].view(np.int32)
    return dict(header=h, desc=desc, x=x, num=num, cur=cur, rowptr=rowptr, order=order, adj=adj, cuv=cuv, cidx=cidx)
