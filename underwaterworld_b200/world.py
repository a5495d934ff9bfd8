"""Host-side mirror of the reference's chunk scheduler, with BATCHED hand-over to the GPU builder.

SURVEY.md §8f-2 / BASELINE config 5.  The reference decides *which* chunks to build in
`World::update_nearby` (src/world.rs:148-234) and then feeds the builder one chunk per frame
(`build_full_step` / `build_step`, src/world.rs:113-145).  This module keeps the selection and the
priority order exactly, and replaces the feeding policy by "one batch per update":

    World.update(sub, camera, builder)      src/world.rs:93-111
      remove_far_way                        src/world.rs:236-247   (all far chunks at once)
      recheck rule (moved > 4 / turned > 0.33 rad / reset)   src/world.rs:5-6,96-103
      update_nearby: 10 x 10 x <=5 window, generation distance, two frusta, GenPrio sort
                                            src/world.rs:148-234, 33-60
      build: pop up to `max_batch` positions in priority order -> ONE ChunkBuilder.build call

Camera / Sub carry only what the scheduler reads (pose, look-at, perspective), following
src/camera.rs:4-62 and src/sub.rs:290-308,376-411,427-448,457-466.  Rendering is out of scope.
Pure numpy (float32 where the reference uses f32); the only GPU work is ChunkBuilder.build.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from .chunk import CHUNK_SIZE, Chunk, ChunkBuilder

f32 = np.float32

# src/world.rs:5-17
RECHECK_NEARBY_DIST = 4.0
RECHECK_NEARBY_ANGLE = 0.33
VIEW_DIST = 4
GENERATION_DIST = 5
STOP_FULL_BUILD = GENERATION_DIST ** 3                       # world.rs:14
KEEP_DIST = 6
MAX_Z = 2
MIN_Z = -2
VIEW_FRUST_FOVY = 55.0
GENERATE_FRUST_FOVY = 90.0
# src/camera.rs:4-6, src/consts.rs:1-2
Z_NEAR = 2.0
Z_FAR = CHUNK_SIZE * (VIEW_DIST + 1)
ASPECT = 800.0 / 600.0
# src/sub.rs:9,14,19-21,29-34
MIDDLE_SPEED = 4.0
MAX_TURN_SPEED = math.pi / 6.0
TARGET_DOWN = 0.6
HORIZONTAL_OFFSET = 7.0
VERTICAL_OFFSET = 6.0
CAMERA_FOLLOW_SPEED = 10.0
START_Y_OFFSET = 0.5 * CHUNK_SIZE
START_Z_OFFSET = 0.75 * CHUNK_SIZE
SUB_MAX_Z = CHUNK_SIZE * 2.0
SUB_MIN_Z = CHUNK_SIZE * -1.5


def _normalize(v):
    n = np.linalg.norm(v)
    return v if n == 0 else v / n


def look_at_rh(eye, target, up) -> np.ndarray:
    """cgmath::Matrix4::look_at_rh (column-vector convention, returned as a 4x4 with M @ [x,y,z,1])."""
    f = _normalize(target - eye)
    s = _normalize(np.cross(f, up))
    u = np.cross(s, f)
    m = np.eye(4, dtype=np.float64)
    m[0, :3], m[1, :3], m[2, :3] = s, u, -f
    m[0, 3], m[1, 3], m[2, 3] = -s.dot(eye), -u.dot(eye), f.dot(eye)
    return m


def perspective(fovy_deg: float, aspect: float, near: float, far: float) -> np.ndarray:
    """cgmath::perspective (OpenGL clip space)."""
    f = 1.0 / math.tan(math.radians(fovy_deg) / 2.0)
    m = np.zeros((4, 4), dtype=np.float64)
    m[0, 0] = f / aspect
    m[1, 1] = f
    m[2, 2] = (far + near) / (near - far)
    m[2, 3] = (2.0 * far * near) / (near - far)
    m[3, 2] = -1.0
    return m


OPENGL_TO_WGPU = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0.5, 0.5], [0, 0, 0, 1]], dtype=np.float64)   # camera.rs:9-14


class Camera:
    """src/camera.rs:17-62 (pose + the generation-frustum matrix)."""

    def __init__(self):
        self.eye = np.array([0.0, 0.0, 0.0])
        self.target = np.array([1.0, 0.0, 0.0])
        self.up = np.array([0.0, 0.0, 1.0])
        self.aspect = ASPECT

    def chunk_generation_frustum_matrix(self, fovy: float) -> np.ndarray:
        view = look_at_rh(self.eye, self.target, self.up)
        proj = perspective(fovy, self.aspect, Z_NEAR, Z_FAR)
        return OPENGL_TO_WGPU @ (proj @ view)


def in_frustum(points: np.ndarray, view_proj: np.ndarray) -> np.ndarray:
    """util::in_frustum (src/util.rs:77-85) on an [n,3] array."""
    p = np.concatenate([points, np.ones((len(points), 1))], axis=1) @ view_proj.T
    w = p[:, 3:4]
    with np.errstate(divide="ignore", invalid="ignore"):
        c = p / w
    return (np.abs(c[:, 0]) <= 1.0) & (np.abs(c[:, 1]) <= 1.0) & (c[:, 2] >= 0.0) & (c[:, 2] <= 1.0)


def _axis_angle(v, axis, angle):
    """Rotate v about a unit axis (cgmath Quaternion::from_axis_angle + rotate_vector)."""
    c, s = math.cos(angle), math.sin(angle)
    return v * c + np.cross(axis, v) * s + axis * axis.dot(v) * (1.0 - c)


class Sub:
    """The submarine's pose integration, scripted instead of keyboard-driven (src/sub.rs:290-308,376-411)."""

    def __init__(self):
        self.pos = np.array([0.0, START_Y_OFFSET, START_Z_OFFSET])
        self.up = np.array([0.0, 0.0, 1.0])
        self.forward = np.array([1.0, 0.0, 0.0])
        self.right = np.array([0.0, 1.0, 0.0])
        self.speed = MIDDLE_SPEED
        self.yaw_speed = 0.0
        self.pitch_speed = 0.0

    def update(self, delta: float):
        mod = min(max(self.speed / MIDDLE_SPEED, 0.0), 1.0)
        pitch_change = self.pitch_speed * delta * mod
        yaw_change = self.yaw_speed * delta * mod
        for name in ("forward", "up", "right"):          # overall = yaw * pitch (roll = 0)
            v = getattr(self, name)
            v = _axis_angle(v, self.right, pitch_change)
            v = _axis_angle(v, self.up, yaw_change)
            setattr(self, name, v)
        self.pos = self.pos + self.forward * self.speed * delta
        self.pos[2] = min(max(self.pos[2], SUB_MIN_Z), SUB_MAX_Z)

    def update_camera(self, camera: Camera, delta: float):
        eye_goal = self.pos - self.forward * HORIZONTAL_OFFSET + self.up * VERTICAL_OFFSET
        camera.eye = camera.eye + (eye_goal - camera.eye) * delta * CAMERA_FOLLOW_SPEED
        target_goal = eye_goal + self.forward - self.up * TARGET_DOWN
        camera.target = camera.target + (target_goal - camera.target) * delta * CAMERA_FOLLOW_SPEED
        camera.up = camera.up + (self.up - camera.up) * delta * CAMERA_FOLLOW_SPEED

    def chunk(self) -> Tuple[int, int, int]:
        return tuple(int(math.floor(v / CHUNK_SIZE)) for v in self.pos)

    def bearing(self):
        return self.forward


_CORNERS = np.array([(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0), (0, 0, 1), (1, 0, 1), (0, 1, 1), (1, 1, 1)], dtype=np.float64)


def nearby_candidates(sub_pos, sub_chunk, camera: Camera):
    """The selection half of World::update_nearby (src/world.rs:152-207) for ALL window positions at once.

    Returns (positions [m,3] int32, dist [m], in_view [m], in_gen [m]) for the chunks within
    GENERATION_DIST, in the reference's enumeration order (x outer, y, z inner)."""
    view_vp = camera.chunk_generation_frustum_matrix(VIEW_FRUST_FOVY)
    gen_vp = camera.chunk_generation_frustum_matrix(GENERATE_FRUST_FOVY)
    start_z = max(sub_chunk[2] - GENERATION_DIST, MIN_Z)
    end_z = min(sub_chunk[2] + GENERATION_DIST, MAX_Z)
    xs = np.arange(-GENERATION_DIST, GENERATION_DIST) + sub_chunk[0]
    ys = np.arange(-GENERATION_DIST, GENERATION_DIST) + sub_chunk[1]
    zs = np.arange(start_z, end_z + 1)
    if len(zs) == 0:
        e = np.zeros(0)
        return np.zeros((0, 3), np.int32), e, e.astype(bool), e.astype(bool)
    g = np.stack(np.meshgrid(xs, ys, zs, indexing="ij"), axis=-1).reshape(-1, 3)
    centers = (g + 0.5) * CHUNK_SIZE
    dist = np.linalg.norm(np.asarray(sub_pos)[None, :] - centers, axis=1)
    keep = dist <= GENERATION_DIST * CHUNK_SIZE
    g, dist = g[keep], dist[keep]
    corners = (g[:, None, :] + _CORNERS[None, :, :]) * CHUNK_SIZE            # [m,8,3]
    flat = corners.reshape(-1, 3)
    gen_hit = in_frustum(flat, gen_vp).reshape(-1, 8)
    view_hit = in_frustum(flat, view_vp).reshape(-1, 8) & gen_hit             # view is only tested inside gen (world.rs:200-205)
    return g.astype(np.int32), dist, view_hit.any(axis=1), gen_hit.any(axis=1)


def gen_prio_order(dist, z_world, in_view, in_gen) -> np.ndarray:
    """Indices in BUILD order: the reference sorts descending by GenPrio::compare and pops from the
    back (src/world.rs:39-60,114,231-233), i.e. builds in ascending GenPrio order:
    in_view first, then in_gen, then smallest dist^2 + z."""
    key = dist * dist + z_world
    return np.lexsort((key, ~in_gen, ~in_view))


@dataclass
class World:
    chunks: Dict[Tuple[int, int, int], Chunk] = field(default_factory=dict)
    chunks_to_render: List[Tuple[int, int, int]] = field(default_factory=list)
    chunks_to_generate: List[Tuple[int, int, int]] = field(default_factory=list)   # in build order
    last_sub_pos: Optional[np.ndarray] = None
    last_sub_bearing: Optional[np.ndarray] = None
    should_full_build: bool = True                            # world.rs:69,83: the start-up phase builds whole chunks

    def get_chunk(self, pos):
        return self.chunks.get(tuple(pos))

    def needs_recheck(self, sub: Sub, sub_reset: bool = False) -> bool:       # world.rs:96-99
        if sub_reset or self.last_sub_pos is None:
            return True
        dist = np.linalg.norm(sub.pos - self.last_sub_pos)
        cosang = float(np.clip(_normalize(sub.bearing()).dot(_normalize(self.last_sub_bearing)), -1.0, 1.0))
        return dist > RECHECK_NEARBY_DIST or math.acos(cosang) > RECHECK_NEARBY_ANGLE

    def update_nearby(self, sub: Sub, camera: Camera):                        # world.rs:148-234
        self.chunks_to_render.clear()
        pos, dist, in_view, in_gen = nearby_candidates(sub.pos, sub.chunk(), camera)
        have = np.array([tuple(p) in self.chunks for p in pos], dtype=bool) if len(pos) else np.zeros(0, bool)
        max_view = VIEW_DIST * CHUNK_SIZE
        for p, d, v in zip(pos[have], dist[have], in_view[have]):
            c = self.chunks[tuple(p)]
            if d < max_view and c.not_blank() and v:
                self.chunks_to_render.append(tuple(int(x) for x in p))
        new = ~have
        order = gen_prio_order(dist[new], pos[new][:, 2].astype(np.float64) * CHUNK_SIZE, in_view[new], in_gen[new])
        self.chunks_to_generate = [tuple(int(x) for x in p) for p in pos[new][order]]

    def remove_far_way(self, sub: Sub):                                       # world.rs:236-247, all at once
        sc = sub.chunk()
        far = [p for p in self.chunks if sum((a - b) ** 2 for a, b in zip(p, sc)) >= KEEP_DIST * KEEP_DIST]
        for p in far:
            del self.chunks[p]

    def build_batch(self, sub: Sub, builder: ChunkBuilder, max_batch: Optional[int] = None) -> int:
        """Batched build_full_step / build_step (world.rs:113-145): pop up to max_batch positions, ONE GPU call.

        Render-list rule, as in the reference: while `should_full_build` is set (start-up, world.rs:105-107) every
        chunk that is not blank is drawn (build_full_step, world.rs:117-119, no distance test); afterwards
        (build_step, world.rs:132-134) only chunks within VIEW_DIST + 1 of the sub's chunk.  The flag is cleared
        once the queue is empty or STOP_FULL_BUILD chunks exist, evaluated after the batch like the reference does
        after every chunk -- with a batch the phase can only end at a batch boundary, so a start-up batch larger than
        what the reference would have built before the switch applies the start-up rule to all of it."""
        take = len(self.chunks_to_generate) if max_batch is None else min(max_batch, len(self.chunks_to_generate))
        if take == 0:
            if self.should_full_build:
                self.should_full_build = False                 # queue empty: world.rs:107
            return 0
        batch_pos, self.chunks_to_generate = self.chunks_to_generate[:take], self.chunks_to_generate[take:]
        batch = builder.build(np.array(batch_pos, dtype=np.int32))
        sc = sub.chunk()
        full = self.should_full_build
        for i, p in enumerate(batch_pos):
            c = Chunk(p)._adopt(batch.chunk(i))
            near = sum((a - b) ** 2 for a, b in zip(p, sc)) <= (VIEW_DIST + 1) ** 2
            if c.not_blank() and (full or near):
                self.chunks_to_render.append(p)
            self.chunks[p] = c
        if full:
            self.should_full_build = not (len(self.chunks_to_generate) == 0 or len(self.chunks) >= STOP_FULL_BUILD)
        return take

    def update(self, sub: Sub, camera: Camera, builder: ChunkBuilder, sub_reset: bool = False,
               max_batch: Optional[int] = None) -> int:                       # world.rs:93-111
        self.remove_far_way(sub)
        if self.needs_recheck(sub, sub_reset):
            self.update_nearby(sub, camera)
            self.last_sub_pos = sub.pos.copy()
            self.last_sub_bearing = sub.bearing().copy()
        return self.build_batch(sub, builder, max_batch)

    # HUD counters, world.rs:251-253
    def generate_count(self):
        return len(self.chunks_to_generate)

    def render_count(self):
        return len(self.chunks_to_render)

    def total_count(self):
        return len(self.chunks)


def scripted_flythrough(frames_straight: int = 600, frames_turn: int = 600, hz: float = 60.0):
    """BASELINE config 5: start (0, 8, 12) heading +x at MIDDLE_SPEED, `frames_straight` frames straight,
    then yaw at MAX_TURN_SPEED for `frames_turn` frames.  Yields (frame, sub, camera) after each step."""
    sub, cam = Sub(), Camera()
    delta = 1.0 / hz
    for _ in range(30):                       # let the camera settle behind the sub (state.rs update order)
        sub.update_camera(cam, delta)
    for frame in range(frames_straight + frames_turn):
        sub.yaw_speed = 0.0 if frame < frames_straight else MAX_TURN_SPEED
        sub.update(delta)
        sub.update_camera(cam, delta)
        yield frame, sub, cam
