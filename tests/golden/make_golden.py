"""Regenerates tests/golden/*.npz from the CPU oracle:  python tests/golden/make_golden.py

PROVENANCE: the reference (FiniteVolumeMethod.jl) cannot be executed in the authoring image — Julia is not
installed and DelaunayTriangulation / OrdinaryDiffEq are not vendored — so these vectors are outputs of
oracle/fvm_oracle.py (the restatement pinned against the reference's own known-answer tests by
tests/test_oracle_golden.py), NOT outputs of the Julia package.  They freeze the oracle: any later change
of the oracle's arithmetic shows up as a fixture mismatch in `pytest -m "not gpu"`, and the CUDA path is
checked against the same frozen numbers in `pytest -m gpu`.  Each file holds the case's inputs (mesh
arrays, u, t) and outputs (du, template A/b, Tsit5 end states, steady solution)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from tests.golden_cases import CASES, oracle_outputs  # noqa: E402


def main():
    for name, make in CASES.items():
        c = make(None)
        tri = c.pair.gtri
        ptr = np.zeros(len(tri.boundary_sections) + 1, dtype=np.int32)
        ptr[1:] = np.cumsum([len(s) for s in tri.boundary_sections])
        data = dict(points=tri.points, triangles=tri.triangles, boundary_ptr=ptr,
                    boundary_nodes=np.concatenate(tri.boundary_sections).astype(np.int32), u=c.u, t=np.float64(c.t))
        data.update(oracle_outputs(c))
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **data)
        print("%-32s %7d bytes  %s" % (name, os.path.getsize(path), ", ".join(sorted(data))))


if __name__ == "__main__":
    main()
