"""CPU restatement of the reference's Gadget-2 (type 1) reader, tests/helper/read_gadget.cuh:15-167
(numpy).  TEST INFRASTRUCTURE: only tests/ may import this.  Follows the reference reader field
by field: header {int npart[6]; double mass[6]; fill to 256 B} between 4-byte markers; blocks
POS, VEL (3 floats x every particle), ID (4 bytes x every particle), MASS (only particles of
types with header mass 0; present only if there is one), U, RHO, HSML (gas only); gas particles
are type 0 and come first in every block."""
import struct

import numpy as np


def read_gadget(path):
    """-> float32 [N_gas, 4] {x, y, z, h} exactly as read_gadget.cuh:69-159 fills h_pos."""
    with open(path, "rb") as f:
        def skip(n_words):          # skip_n, :36-42
            f.seek(4 * n_words, 1)
        skip(1)                     # read_gadget_header, :58-67
        npart = struct.unpack("<6i", f.read(24))
        mass = struct.unpack("<6d", f.read(48))
        f.seek(256 - 24 - 48, 1)
        skip(1)
        n_gas = npart[0]
        if n_gas == 0:
            raise RuntimeError("Gadget file %s has no gas particles!" % path)     # :85-90
        n_total = sum(npart)
        n_withmass = sum(n for n, m in zip(npart, mass) if m == 0)                  # :99-101
        out = np.empty((n_gas, 4), np.float32)
        skip(1)
        out[:, :3] = np.frombuffer(f.read(12 * n_gas), "<f4").reshape(-1, 3)        # :103-108
        skip(3 * (n_total - n_gas))
        skip(1)
        skip(1 + 3 * n_total + 1)                                                   # VEL :123
        skip(1 + n_total + 1)                                                       # ID  :126
        if n_withmass > 0:                                                          # :131-141
            skip(1 + n_withmass + 1)
        skip(1 + n_gas + 1)                                                         # U   :147-149
        skip(1 + n_gas + 1)                                                         # RHO :152-154
        skip(1)
        out[:, 3] = np.frombuffer(f.read(4 * n_gas), "<f4")                         # :157-161
        return out


def write_gadget(path, spheres, n_other=0, other_has_mass_block=False, vel=None):
    """A file read_gadget.cuh loads: gas = spheres [N,4] {x,y,z,h}; n_other type-1 particles."""
    s = np.ascontiguousarray(spheres, np.float32)
    n_gas, n_total = len(s), len(s) + n_other
    rng = np.random.default_rng(7)

    def block(a):
        b = np.ascontiguousarray(a).tobytes()
        return struct.pack("<i", len(b)) + b + struct.pack("<i", len(b))
    hdr = struct.pack("<6i", n_gas, n_other, 0, 0, 0, 0) + \
        struct.pack("<6d", 1.0, 0.0 if other_has_mass_block else 2.0, 0, 0, 0, 0)
    hdr += b"\0" * (256 - len(hdr))
    pos = np.concatenate([s[:, :3], rng.random((n_other, 3), dtype=np.float32)])
    with open(path, "wb") as f:
        f.write(block(np.frombuffer(hdr, np.uint8)))
        f.write(block(pos))
        f.write(block(rng.random((n_total, 3), dtype=np.float32) if vel is None else vel))
        f.write(block(np.arange(n_total, dtype=np.int32)))
        if other_has_mass_block and n_other:
            f.write(block(np.full(n_other, 3.0, np.float32)))
        f.write(block(rng.random(n_gas, dtype=np.float32)))      # U
        f.write(block(rng.random(n_gas, dtype=np.float32)))      # RHO
        f.write(block(s[:, 3]))                                  # HSML
