"""The one constant of the reference's helperFunctions module the hot path needs
(helperFunctions.py:20).  The reference's own helperFunctions stays in use for everything else."""
eps = 1e-6
