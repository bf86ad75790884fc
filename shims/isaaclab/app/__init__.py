"""AppLauncher stand-in: there is no Omniverse Kit to boot; keeps the CLI flags the reference passes."""
import argparse


class _App:
    def is_running(self):
        return True

    def close(self):
        pass


class AppLauncher:
    def __init__(self, launcher_args=None, **kwargs):
        self._args = launcher_args
        self.app = _App()

    @staticmethod
    def add_app_launcher_args(parser: argparse.ArgumentParser) -> None:
        g = parser.add_argument_group("app_launcher arguments (stand-in)")
        g.add_argument("--headless", action="store_true", default=False)
        g.add_argument("--livestream", type=int, default=-1)
        g.add_argument("--enable_cameras", action="store_true", default=False)
        g.add_argument("--device", type=str, default="cuda:0")
        g.add_argument("--verbose", action="store_true", default=False)
        g.add_argument("--experience", type=str, default="")
        g.add_argument("--kit_args", type=str, default="")
