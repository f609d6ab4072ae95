"""Decodes the reference's own data sets into small JSON fixtures.

Run once in the authoring container (needs /root/reference, which does not
exist on the GPU box):   python tests/golden/make_fixtures.py

  /root/reference/data/poisson_series.rda -> poisson_series.json  (100 ints)
  /root/reference/data/exchange.rda       -> exchange.json        (945 doubles)

.rda = gzip/bzip2 stream "RDX2\\n" + XDR serialisation: a pairlist whose CAR is
the vector (INTSXP 13 / REALSXP 14), possibly with attributes (ts).  The reader
below scans for the first vector header of the wanted type and length.
"""
import bz2
import gzip
import json
import os
import struct

REF = "/root/reference/data"
HERE = os.path.dirname(os.path.abspath(__file__))


def read_rda(path):
    raw = open(path, "rb").read()
    if raw[:2] == b"\x1f\x8b":
        return gzip.decompress(raw)
    if raw[:3] == b"BZh":
        return bz2.decompress(raw)
    return raw


def find_vector(buf, sexptype, length):
    """first XDR header (flags & 0xff == sexptype) followed by `length`"""
    for off in range(5, len(buf) - 8):
        flags, ln = struct.unpack(">II", buf[off:off + 8])
        if (flags & 0xff) == sexptype and (flags >> 16) == 0 and ln == length:
            body = off + 8
            if sexptype == 13:
                return list(struct.unpack(">%di" % length, buf[body:body + 4 * length]))
            return list(struct.unpack(">%dd" % length, buf[body:body + 8 * length]))
    raise ValueError("vector not found")


def main():
    pois = find_vector(read_rda(os.path.join(REF, "poisson_series.rda")), 13, 100)
    assert pois[:12] == [2, 0, 0, 2, 0, 1, 3, 0, 1, 3, 1, 0] and pois[-3:] == [18, 16, 11]
    json.dump({"source": "data/poisson_series.rda", "y": pois},
              open(os.path.join(HERE, "poisson_series.json"), "w"))
    exch = find_vector(read_rda(os.path.join(REF, "exchange.rda")), 14, 945)
    assert abs(exch[0] - (-0.3555316)) < 1e-6 and abs(exch[1] - 1.4254090) < 1e-6
    json.dump({"source": "data/exchange.rda", "y": exch},
              open(os.path.join(HERE, "exchange.json"), "w"))
    print("poisson_series", len(pois), "exchange", len(exch))


if __name__ == "__main__":
    main()
