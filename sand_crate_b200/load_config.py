"""YAML -> config objects, same schema as the reference's `src/crate/load_config.py:7-46`
(`world: {coefficients, particle_sources, rigid_bodies}`, `playback: {...}`), so `config/*.yaml` files written for
SandCrate load unchanged."""
from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path
from typing import Any

import yaml


@dataclass
class WorldConfig:
    rigid_bodies: list = field(default_factory=list)
    particle_sources: list = field(default_factory=list)
    coefficients: dict = field(default_factory=dict)


@dataclass
class PlaybackConfig:
    save_recording: bool = False
    ticks_to_record: int = 0
    recording_output_dir_path: Path = Path(".")
    screen_x: int = 1000
    screen_y: int = 1000


@dataclass
class Config:
    world_config: WorldConfig
    playback_config: PlaybackConfig


def config_from_dict(raw: dict[str, Any]) -> Config:
    world = raw["world"]
    wc = WorldConfig(
        rigid_bodies=world.get("rigid_bodies", []) or [],
        particle_sources=world.get("particle_sources") or [],
        coefficients=world.get("coefficients") or {},
    )
    pb = raw.get("playback") or {}
    pc = PlaybackConfig(
        save_recording=pb.get("save_recording", False),
        ticks_to_record=pb.get("ticks_to_record", 0),
        recording_output_dir_path=Path(pb.get("recording_output_dir_path", ".")),
        screen_x=pb.get("screen_x", 1000),
        screen_y=pb.get("screen_y", 1000),
    )
    return Config(world_config=wc, playback_config=pc)


def load_config(config_file_path) -> Config:
    with open(config_file_path, "r") as f:
        return config_from_dict(yaml.safe_load(f))
