"""Drop-in for the reference's ``pose_video`` package (hot-path modules only)."""
