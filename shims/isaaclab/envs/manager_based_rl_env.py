from . import ManagerBasedRLEnv, ManagerBasedRLEnvCfg  # noqa: F401  (curriculums.py:3 imports from this module)
