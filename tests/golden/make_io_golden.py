"""Golden I/O fixtures (run in the build container, where /root/reference exists):

  io_u16.his + io_u16_frames.npy   a small multi-frame u16 HIS file and what the REFERENCE's his::load
                                   (src/his.cpp, compiled unmodified into oracle/_ref) decoded from it
  io_ref.ddbvf + io_ref_volume.npy a DDBVF file written by the REFERENCE's ddbvf::create/write (src/ddbvf.cpp)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import formats  # noqa: E402

rng = np.random.default_rng(20261018)
frames = rng.integers(0, 65535, size=(3, 6, 8), endpoint=True).astype(np.uint16)
formats.write_his(os.path.join(HERE, "io_u16.his"), frames, 4, image_header_size=32, ulx=2, uly=1)
np.save(os.path.join(HERE, "io_u16_frames.npy"), formats.RefIO().his_load(os.path.join(HERE, "io_u16.his")))

vol = rng.standard_normal((4, 3, 5)).astype(np.float32)
formats.RefIO().ddbvf_create_write(os.path.join(HERE, "io_ref"), (5, 3, 4), vol, 0)
np.save(os.path.join(HERE, "io_ref_volume.npy"), vol)
print("written")
