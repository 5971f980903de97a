"""Reader for the records written by oracle/ref_dump.cpp (and for the .npz fixtures made from them).

TEST INFRASTRUCTURE ONLY (see oracle/dmrg_oracle.py header)."""
from __future__ import annotations

import struct

import numpy as np

from . import dmrg_oracle as O


def read_records(path) -> dict:
    """{ u32 name_len, name, u8 dtype (0=i32,1=f64), u32 ndim, u64 dims[], data } repeated; or an .npz of the same."""
    if str(path).endswith(".npz"):
        with np.load(path) as z:
            return {k: z[k] for k in z.files}
    out = {}
    with open(path, "rb") as f:
        buf = f.read()
    p = 0
    while p < len(buf):
        (n,) = struct.unpack_from("<I", buf, p); p += 4
        name = buf[p:p + n].decode(); p += n
        dtype = buf[p]; p += 1
        (nd,) = struct.unpack_from("<I", buf, p); p += 4
        dims = struct.unpack_from("<%dQ" % nd, buf, p); p += 8 * nd
        cnt = int(np.prod(dims)) if nd else 1
        if dtype == 0:
            arr = np.frombuffer(buf, dtype="<i4", count=cnt, offset=p); p += 4 * cnt
        else:
            arr = np.frombuffer(buf, dtype="<f8", count=cnt, offset=p); p += 8 * cnt
        out[name] = arr.reshape(dims).copy()
    return out


def write_records(path, rec: dict) -> None:
    """Inverse of read_records (raw format): lets C++ test drivers consume the .npz fixtures."""
    with open(path, "wb") as f:
        for name, arr in rec.items():
            a = np.asarray(arr)
            isint = a.dtype.kind in "iub"
            a = np.ascontiguousarray(a, dtype="<i4" if isint else "<f8")
            nb = name.encode()
            f.write(struct.pack("<I", len(nb))); f.write(nb)
            f.write(struct.pack("<B", 0 if isint else 1))
            f.write(struct.pack("<I", a.ndim))
            f.write(struct.pack("<%dQ" % a.ndim, *a.shape))
            f.write(a.tobytes())


def block_from(rec: dict, prefix: str) -> O.Block:
    q = rec[prefix + "q"].astype(np.int64).reshape(-1, 3)
    dims = rec[prefix + "dims"].astype(np.int64)
    blk = O.Block(q=q, dims=dims, sites=tuple(int(s) for s in rec[prefix + "sites"]), loop=bool(rec[prefix + "flags"][0]))
    nops = int(rec[prefix + "nops"][0])
    for m in range(nops):
        meta = rec["%sop%d.meta" % (prefix, m)]
        allowed = rec["%sop%d.allowed" % (prefix, m)].astype(bool)
        data = rec["%sop%d.data" % (prefix, m)]
        norb = int(meta[2])
        orbs = tuple(int(x) for x in meta[3:3 + norb])
        blocks, off = {}, 0
        for i in range(allowed.shape[0]):
            for j in range(allowed.shape[1]):
                if allowed[i, j]:
                    n = int(dims[i]) * int(dims[j])
                    blocks[(i, j)] = data[off:off + n].reshape(int(dims[i]), int(dims[j]))
                    off += n
        assert off == data.size
        blk.ops.append(O.Op(optype=int(meta[0]), orbs=orbs, comp=int(meta[5]), dq=(int(meta[6]), int(meta[7]), int(meta[8])),
                            fermion=bool(meta[9]), allowed=allowed, blocks=blocks))
    return blk


def big_from(rec: dict) -> O.Big:
    L, R = block_from(rec, "L."), block_from(rec, "R.")
    meta = rec["meta"]
    return O.Big(left=L, right=R, psi_dq=tuple(int(x) for x in rec["psi_dq"]), core_energy=float(rec["meta_f"][3]),
                 hubbard=(int(meta[7]) == O.HUBBARD_HAM))


def rotation_from(rec: dict):
    shape = rec["rot.shape"].reshape(-1, 2)
    data = rec["rot.data"]
    out, off = [], 0
    for nr, nc in shape:
        nr, nc = int(nr), int(nc)
        if nc == 0:
            out.append(np.zeros((nr, 0)))
        else:
            out.append(data[off:off + nr * nc].reshape(nr, nc)); off += nr * nc
    return out
