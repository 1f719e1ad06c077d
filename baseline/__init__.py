"""The reference (ulrikisdahl/MuZero-Breakout) as a baseline arm.  `baseline/_ref/` is a git-ignored copy of the unmodified
checkout made by `__graft_entry__.build()` in the build container (the reference is pure Python: there is nothing to
compile, and /root/reference does not exist on the GPU box); `baseline/ref.py` makes it importable."""
