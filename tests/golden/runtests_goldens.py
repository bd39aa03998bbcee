"""Golden values transcribed from the reference's own test-suite (test/runtests.jl), fixture
demo/pincell.json -> tests/golden/pincell.npz (make_pincell_fixture.py).  These are the only
known-answer vectors the reference holds for the trace! -> segmentize! path."""

# test/runtests.jl:10-27  TrackGenerator(model, 8, 0.02) (default Vacuum BCs)
MAIN = dict(
    n_azim=8, delta=0.02, n_total_tracks=420, n_tracks_x=[31, 74, 74, 31], n_tracks_y=[74, 31, 31, 74],
    n_tracks=[105, 105, 105, 105], nazim2=4, nazim4=2, delta_eff=0.01994243696980254,
    phis=[0.39670866289121387, 1.1740876639036828, 1.9675049896861103, 2.7448839906985794],
)

# test/runtests.jl:48  bcs = (top=Vacuum, bottom=Reflective, left=Reflective, right=Vacuum)
REFLECTION_BCS = dict(top="Vacuum", bottom="Reflective", left="Reflective", right="Vacuum")

# uid: (bc_fwd, bc_bwd, next_fwd_uid, next_bwd_uid, dir_fwd, dir_bwd)  V/R = Vacuum/Reflective, F/B = Forward/Backward
# test/runtests.jl:49-162  TrackGenerator(model, 4, 0.8; bcs)
LINKS_4 = {
    1: ("V", "R", 7, 6, "F", "F"), 2: ("V", "R", 8, 5, "F", "F"), 3: ("V", "R", 8, 5, "B", "B"),
    4: ("V", "R", 7, 6, "B", "B"), 5: ("R", "R", 3, 2, "F", "F"), 6: ("R", "R", 4, 1, "F", "F"),
    7: ("V", "V", 4, 1, "B", "B"), 8: ("V", "V", 3, 2, "B", "B"),
}
# test/runtests.jl:165-334  TrackGenerator(model, 8, 0.8; bcs)
LINKS_8 = {
    1: ("V", "R", 11, 10, "F", "F"), 2: ("V", "R", 12, 10, "F", "B"), 3: ("V", "R", 12, 11, "B", "B"),
    4: ("V", "R", 9, 8, "F", "F"), 5: ("V", "R", 9, 7, "B", "F"), 6: ("V", "R", 8, 7, "B", "B"),
    7: ("R", "R", 6, 5, "F", "F"), 8: ("V", "R", 6, 4, "B", "F"), 9: ("V", "V", 5, 4, "B", "B"),
    10: ("R", "R", 2, 1, "F", "F"), 11: ("R", "V", 3, 1, "F", "B"), 12: ("V", "V", 3, 2, "B", "B"),
}
BC_CODE = {"V": 0, "R": 1, "P": 2}
DIR_CODE = {"F": 0, "B": 1}
