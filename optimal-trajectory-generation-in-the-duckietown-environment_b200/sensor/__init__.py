"""Obstacles and the cone proximity sensor (same exports as lib/sensor/__init__.py)."""
from .obstacles import Obstacle, StaticObstacle, MovingObstacle, ObstacleTrajectory
from .proximity_sensor import Sensor, ProximitySensor

__all__ = ["Sensor", "ProximitySensor", "Obstacle", "StaticObstacle", "MovingObstacle", "ObstacleTrajectory"]
