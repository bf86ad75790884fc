"""gymnasium.core stand-in: Env / Wrapper / ActionWrapper with attribute forwarding (what ClipAction relies on)."""
from typing import TypeVar

ActType = TypeVar("ActType")
ObsType = TypeVar("ObsType")
WrapperActType = TypeVar("WrapperActType")
WrapperObsType = TypeVar("WrapperObsType")
RenderFrame = TypeVar("RenderFrame")


class Env:
    metadata: dict = {"render_modes": []}
    render_mode = None
    spec = None

    @property
    def unwrapped(self):
        return self

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith("_") and name != "_np_random":
            raise AttributeError(f"accessing private attribute '{name}' is prohibited")
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    @property
    def action_space(self):
        return self.__dict__.get("_action_space") or self.env.action_space

    @action_space.setter
    def action_space(self, space):
        self.__dict__["_action_space"] = space

    @property
    def observation_space(self):
        return self.__dict__.get("_observation_space") or self.env.observation_space

    @observation_space.setter
    def observation_space(self, space):
        self.__dict__["_observation_space"] = space

    def step(self, action):
        return self.env.step(action)

    def reset(self, *, seed=None, options=None):
        return self.env.reset(seed=seed, options=options)

    def render(self):
        return self.env.render()

    def close(self):
        return self.env.close()


class ActionWrapper(Wrapper):
    def step(self, action):
        return self.env.step(self.action(action))

    def action(self, action):
        raise NotImplementedError


class ObservationWrapper(Wrapper):
    def step(self, action):
        obs, rew, term, trunc, info = self.env.step(action)
        return self.observation(obs), rew, term, trunc, info

    def observation(self, observation):
        raise NotImplementedError


class RewardWrapper(Wrapper):
    def step(self, action):
        obs, rew, term, trunc, info = self.env.step(action)
        return obs, self.reward(rew), term, trunc, info

    def reward(self, reward):
        raise NotImplementedError
