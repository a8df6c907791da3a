"""
CPU oracle for the aMOF hot path -- TEST INFRASTRUCTURE, never imported by ``amof_b200``.

``oracle.c_oracle``   ctypes binding of liboracle.so (oracle/amof_oracle.c)
``oracle.np_oracle``  independent numpy twin for small cases
``oracle.ref_classes`` line-by-line restatement of the reference's driver code
                      (amof/rdf.py, cn.py, bad.py, msd.py) on top of the C oracle

Parity status: **parity unpinned** (see amof_oracle.c header and DESIGN.md).
"""
