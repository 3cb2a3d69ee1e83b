// Intentionally empty: the beam-field variants are always compiled (track_fast.cu /
// track_strict.cu with -DXLB_BEAMFIELDS=1).  Kept so older build trees link.
