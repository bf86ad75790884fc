"""Stand-in for the import surface of IsaacLab v2.0.2 that the UNMODIFIED WheeledLab sources touch
(SURVEY.md 8b).  NOT a re-implementation of IsaacLab: configuration classes carry data only, and the
``mdp`` accessors read tensors off whatever env object they are given.  Used (a) by tests/golden/make_golden.py
to execute the reference's own term functions in this container and (b) as the shim that lets the reference's
task packages import against wheeledlab_b200.  Behaviour recalled from upstream is tagged [UPSTREAM-RECALL].
"""
__version__ = "2.0.2-standin"
