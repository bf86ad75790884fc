"""Import-time stand-in for matplotlib (wheeledlab_tasks/visual/utils/traversability_utils.py:2 imports pyplot for a debug
plot that the env path never calls).  Used only when the real package is absent."""
