"""Writes tests/golden/reference_tests.json: the known-answer vectors the reference's own
test-suite holds for the hot path's index maps and coefficients, transcribed by hand from
/root/reference/test/*.jl (file:line recorded per entry).  The reference is Julia and cannot
be executed in this image, so these are transcriptions of its assertions, not generated
outputs; this script only serialises them so that the tests read one file."""
import json
import math
import os

inf = "inf"
G = {
    "source": "facebookresearch/Khronos.jl test/ (v0.2.0), transcribed",
    # test/test_grid_volume.jl:13-118 — 80^3 grid (cell 4.0, resolution 20), default Float64 backend
    "grid_volume": {
        "cite": "test/test_grid_volume.jl:13-118",
        "sim": {"cell_size": [4.0, 4.0, 4.0], "cell_center": [0.0, 0.0, 0.0], "resolution": 20, "dtype": "f64"},
        "cases": [
            {"size": [0, 0, 0], "comp": "Ex", "N": [2, 1, 1], "start": [40, 41, 41], "end": [41, 41, 41]},
            {"size": [0, 0, 0], "comp": "Ey", "N": [1, 2, 1], "start": [41, 40, 41], "end": [41, 41, 41]},
            {"size": [0, 0, 0], "comp": "Ez", "N": [1, 1, 2], "start": [41, 41, 40], "end": [41, 41, 41]},
            {"size": [0, inf, 0], "comp": "Ex", "N": [2, 81, 1], "start": [40, 1, 41], "end": [41, 81, 41]},
            {"size": [0, inf, 0], "comp": "Ey", "N": [1, 80, 1], "start": [41, 1, 41], "end": [41, 80, 41]},
            {"size": [0, inf, 0], "comp": "Ez", "N": [1, 81, 2], "start": [41, 1, 40], "end": [41, 81, 41]},
            {"size": [0, inf, inf], "comp": "Ex", "N": [2, 81, 81], "start": [40, 1, 1], "end": [41, 81, 81]},
            {"size": [0, inf, inf], "comp": "Ey", "N": [1, 80, 81], "start": [41, 1, 1], "end": [41, 80, 81]},
            {"size": [0, inf, inf], "comp": "Ez", "N": [1, 81, 80], "start": [41, 1, 1], "end": [41, 81, 80]},
        ],
    },
    # test/test_sources.jl:12-83 — 100^3 grid with PML 1.0; off-grid point touches 8 voxels, a line 4*N
    "source_footprint": {
        "cite": "test/test_sources.jl:12-83",
        "sim": {"cell_size": [10.0, 10.0, 10.0], "cell_center": [0.0, 0.0, 0.0], "resolution": 10, "dtype": "f64"},
        "center": [0.023, 0.784, 0.631],
        "point_voxels": 8,
        "line_voxels_factor": 4,
    },
    # test/test_chunking.jl:640-797 — 100^3 grid, PML 1.0 on every side
    "pml_grid": {
        "cite": "test/test_chunking.jl:640-797",
        "sim": {"cell_size": [10.0, 10.0, 10.0], "cell_center": [0.0, 0.0, 0.0], "resolution": 10, "dtype": "f64"},
        "boundaries_all": [[1.0, 1.0], [1.0, 1.0], [1.0, 1.0]], "regions_all": 27,
        "boundaries_x_only": [[1.0, 1.0], [0.0, 0.0], [0.0, 0.0]], "regions_x_only": 3,
        "class_counts": {"interior": 1, "face": 6, "edge": 12, "corner": 8},
        "adjacencies": 54, "interior_neighbours": 6,
    },
    # test/test_chunking.jl:799-858 — which aux arrays exist (allocation rule Chunking.jl:1091-1122)
    "aux_pattern": {
        "cite": "test/test_chunking.jl:799-858",
        "interior": {"pml": [0, 0, 0], "allocated": []},
        "face_x": {"pml": [1, 0, 0], "allocated_includes": ["WBx"], "absent": ["UBx", "WBy", "WBz"]},
    },
    # test/test_chunking.jl:860-908 — sigma arrays of a face-x chunk are zero on the non-PML axes
    "chunk_sigma": {"cite": "test/test_chunking.jl:860-908", "face_x_zero_axes": [1, 2]},
    # test/test_dispersive.jl:22-40
    "ade": {
        "cite": "test/test_dispersive.jl:22-40", "dt": 0.01,
        "lorentz": {"omega_0": 1.0, "gamma": 0.1, "sigma": 2.0,
                    "gamma1": 1.0 - 0.1 * math.pi * 0.01, "gamma1_inv": 1.0 / (1.0 + 0.1 * math.pi * 0.01),
                    "omega0_dt_sq": (2 * math.pi * 1.0 * 0.01) ** 2, "is_drude": False},
        "drude": {"gamma": 0.5, "sigma": 3.0, "omega0_dt_sq": 0.0, "drude_coeff": 0.5 * 2 * math.pi * 0.01 ** 2,
                  "is_drude": True},
    },
    # test/test_interpolation.jl:42-92 — interpolation weights of a volume sum to its measure / to 1
    "interpolation": {"cite": "test/test_interpolation.jl:42-92", "tolerance": 1e-12},
    # test/test_absorber.jl:55-81,147-160 — ramp monotone towards the boundary, zero in the interior
    "absorber": {"cite": "test/test_absorber.jl:22-82,123-161"},
}
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_tests.json")
with open(out, "w") as f:
    json.dump(G, f, indent=1)
print("wrote", out)
