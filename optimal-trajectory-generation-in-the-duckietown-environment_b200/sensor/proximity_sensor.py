"""Cone proximity sensor (lib/sensor/proximity_sensor.py:13-89)."""
from __future__ import annotations

import numpy as np


class Sensor:
    pass


class ProximitySensor(Sensor):
    """Reports the obstacles inside a cone of length `range` and half-aperture `aperture` ahead of the robot."""

    def __init__(self, range: float, aperture: float, *args, **kwargs):
        self.range = range
        self.aperture = aperture
        self.obstacles = []
        if kwargs:
            self.__dict__.update(kwargs)

    def attach(self, robot):
        self.robot = robot

    def set_obstacle(self, obs_lst):
        self.obstacles = obs_lst

    def active_mask(self, robot_pose) -> np.ndarray:
        """Bool mask over `self.obstacles` (list order) of what the cone sees from `robot_pose`."""
        p, th = np.asarray(robot_pose[:2], dtype=np.float64), robot_pose[2]
        heading = np.array([np.cos(th), np.sin(th)])
        mask = np.zeros(len(self.obstacles), dtype=bool)
        for i, ob in enumerate(self.obstacles):
            rel = ob.position - p
            along = np.dot(rel, heading)
            if along < 0. or along >= self.range:
                continue
            angle = np.arctan2(abs(heading[0] * rel[1] - heading[1] * rel[0]), np.dot(heading, rel))
            if np.abs(angle) > self.aperture:
                continue
            mask[i] = True
        return mask

    def sense(self, *args, **kwargs):
        """Obstacles inside the cone.  A positional pose is ignored exactly like in the reference (:53-57):
        only the keyword `rpose=` overrides the attached robot's pose."""
        pose = kwargs['rpose'] if 'rpose' in kwargs else self.robot.pose()
        mask = self.active_mask(pose)
        return [ob for ob, m in zip(self.obstacles, mask) if m]

    def step_obstacle(self, time: float):
        for ob in self.obstacles:
            ob.step(time)
        self.obstacles = list(self.obstacles)
