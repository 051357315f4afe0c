"""CPU oracle for the SVOL hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``svol_b200/`` imports this package.  See ``svol_oracle.py`` for the numpy
restatement of the reference's forward / matcher / criterion and ``lsap.py`` / ``lsap.c``
for the restated rectangular assignment solver.
"""
