"""On-disk template store of the reference: webdataset-style tar shards holding, per mesh
and view, ``{key}_{k}.rgb.png`` (RGB u8) and ``{key}_{k}.depth.png`` (16-bit, millimetres), 10 meshes per shard
``shard-%06d.tar``; ``key`` is the mesh id with underscores removed.

* :class:`TemplateShardWriter` / :func:`render_templates` -- the writer side of reference ``scripts/render_templates.py:49-72``
  (webdataset.ShardWriter with PNG encoding of numpy arrays), fed by the device rasteriser.
* :class:`WebTemplateDataset` -- the reader of reference ``src/dataloader/template.py:26-99`` with the same constructor,
  ``__getitem__`` / ``get_template_by_name`` and return schema, so shards rendered by either side are interchangeable.

This is host I/O (tar + PNG through PIL, like the reference); nothing here is on the per-proposal hot path -- with the
device rasteriser the estimators render templates online (``DinoPoseEstimator.forward_mesh``) and the store is only
needed for drop-in compatibility with pre-rendered data."""
from __future__ import annotations

import io
import os
import tarfile
import time
from pathlib import Path

import numpy as np
import torch

from .bbox_utils import CropResizePad
from .utils import mask_to_bbox

MESHES_PER_SHARD = 10          # template.py:52 (idx // 10), render_templates.py:41-42
VIEWS = 600


def _png_bytes(arr: np.ndarray) -> bytes:
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(arr).save(buf, format="PNG")
    return buf.getvalue()


class TemplateShardWriter:
    """``wds.ShardWriter(pattern, start_shard=...)`` reduced to what render_templates.py uses: consecutive tar members
    ``<__key__>.<ext>``; a new shard every ``meshes_per_shard`` meshes."""

    def __init__(self, shards_dir, start_shard: int = 0, meshes_per_shard: int = MESHES_PER_SHARD):
        self.dir = Path(shards_dir)
        self.dir.mkdir(parents=True, exist_ok=True)
        self.shard = start_shard
        self.per_shard = meshes_per_shard
        self._tar = None
        self._count = 0

    def _open(self):
        self._tar = tarfile.open(self.dir / f"shard-{self.shard:06d}.tar", "w")
        self._count = 0

    def _add(self, name: str, data: bytes):
        ti = tarfile.TarInfo(name)
        ti.size = len(data)
        ti.mtime = time.time()
        ti.mode = 0o444
        self._tar.addfile(ti, io.BytesIO(data))

    def write_mesh(self, mesh_id: str, rgb: np.ndarray, depth: np.ndarray):
        """rgb (V,H,W,3) u8, depth (V,H,W) fp32 metres -> 2V tar members (render_templates.py:66-72: depth is stored as
        ``(depth * 1000).astype(uint16)``, i.e. truncated millimetres)."""
        if self._tar is None:
            self._open()
        elif self._count == self.per_shard:
            self.close()
            self.shard += 1
            self._open()
        key = mesh_id.replace("_", "")
        for i in range(len(rgb)):
            self._add(f"{key}_{i}.rgb.png", _png_bytes(np.ascontiguousarray(rgb[i], dtype=np.uint8)))
            self._add(f"{key}_{i}.depth.png", _png_bytes((np.asarray(depth[i]) * 1000).astype(np.uint16)))
        self._count += 1

    def close(self):
        if self._tar is not None:
            self._tar.close()
            self._tar = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def render_templates(meshes: dict, shards_dir, n_poses: int = VIEWS, resolution: int = 420, start_shard: int = 0,
                     scale: float = 0.25, device="cuda"):
    """render_templates.py:49-72 with the device rasteriser: every mesh (id -> mesh, at unit normalisation) is scaled by
    0.25, rendered from the n_poses template views with SKIP_CULL_FACES and appended to the shards."""
    from .retrieval.renderer import MeshRenderer
    from .utils import as_mesh
    renderer = MeshRenderer(n_poses, resolution=resolution, device=device)
    with TemplateShardWriter(shards_dir, start_shard) as w:
        for mesh_id, mesh in meshes.items():
            m = as_mesh(mesh).copy().apply_scale(scale)
            rgb, depth = renderer.render_device(m, None, cull_faces=False)
            w.write_mesh(mesh_id, rgb.cpu().numpy(), depth.cpu().numpy())


def collate_fn(batch):
    batch = [b for b in batch if b["templates"] is not None]
    if len(batch) == 0:
        return None
    return {"templates": torch.cat([b["templates"] for b in batch]), "model_name": [b["model_name"] for b in batch],
            "tar_file": [b["tar_file"] for b in batch]}


class WebTemplateDataset(torch.utils.data.Dataset):
    """Same surface as the reference class.  ``filelist_path`` is a CSV with a ``model_name`` column (row i lives in
    shard i // 10); ``n_views`` (600 in the reference, hard-coded there) is configurable for small test stores."""

    def __init__(self, wds_dir: str, filelist_path: str, resolution: int = 420, bbox_extend: int = 0, crop: bool = True,
                 n_views: int = VIEWS):
        import pandas as pd
        self.wds_dir = Path(wds_dir).resolve()
        self.frame_index = pd.read_csv(Path(filelist_path).resolve(), dtype=str)["model_name"]
        self.rgb_proposal_processor = CropResizePad(resolution, (420, 420), bbox_extend=bbox_extend)
        self.crop = crop
        self.n_views = n_views
        self.frame_index = self.frame_index.str.replace("_", "")
        self._index = {}     # shard path -> {member name: (data offset, size)}; the reference pickles a TarInfo dict beside
                             # every shard (template.py:54-61), here the index lives in memory for the process lifetime

    def __len__(self):
        return len(self.frame_index)

    def get_template_by_name(self, model_name):
        idx = self.frame_index[self.frame_index == model_name].index[0]
        return self.__getitem__(idx)

    def _decode(self, idx: int):
        """-> rgb u8 (V,H,W,3), depth (V,H,W) integer millimetres, model_name, tar name.  The 2 x n_views PNG members
        are read with raw seeks (one pass over the shard file) and decoded on a thread pool -- PIL releases the GIL
        inside the decoder; the reference decodes its 1200 images per mesh one after the other (template.py:65-72)."""
        from concurrent.futures import ThreadPoolExecutor
        from PIL import Image
        tar_path = self.wds_dir / f"shard-{idx // MESHES_PER_SHARD:06d}.tar"
        model_name = self.frame_index[idx].replace("_", "")
        key = tar_path.as_posix()
        if key not in self._index:
            with tarfile.open(key) as tar:
                self._index[key] = {m.name: (m.offset_data, m.size) for m in tar.getmembers()}
        members = self._index[key]
        blobs = []
        with open(key, "rb") as f:
            for k in range(self.n_views):
                for kind in ("rgb", "depth"):
                    off, size = members[f"{model_name}_{k}.{kind}.png"]
                    f.seek(off)
                    blobs.append(f.read(size))

        def decode(i):
            img = Image.open(io.BytesIO(blobs[i]))
            return np.array(img.convert("RGB")) if i % 2 == 0 else np.array(img)

        with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as pool:
            arrays = list(pool.map(decode, range(len(blobs))))
        if not arrays:
            return None, None, model_name, tar_path.name
        return np.stack(arrays[0::2]), np.stack(arrays[1::2]), model_name, tar_path.name

    def __getitem__(self, idx: int):
        rgb, dep, model_name, tar_name = self._decode(idx)
        if rgb is None:
            return {"templates": None, "masks": None, "depths": None, "bboxes": None, "model_name": model_name,
                    "tar_file": tar_name}
        V = rgb.shape[0]
        # template.py:73-80 for all views at once, in the reference's arithmetic: float64 division, then .float()
        # (chunked: 600 views x 420^2 x 3 in float64 would be 2.5 GB of temporaries)
        templates = torch.empty(V, 3, rgb.shape[1], rgb.shape[2], dtype=torch.float32)
        depths = torch.empty(V, dep.shape[1], dep.shape[2], dtype=torch.float32)
        for i in range(0, V, 32):
            templates[i:i + 32] = torch.from_numpy(rgb[i:i + 32] / 255).float().permute(0, 3, 1, 2)
            depths[i:i + 32] = torch.from_numpy(dep[i:i + 32] / 1000).float()
        masks = depths > 0
        small = masks.flatten(1).sum(1) < 100
        masks[small, 105:315, 105:315] = True          # template.py:74-76
        # mask_to_bbox per view = first / last set column and row
        cols, rows = masks.any(1).numpy(), masks.any(2).numpy()
        W, H = cols.shape[1], rows.shape[1]
        bboxes = torch.from_numpy(np.stack([cols.argmax(1), rows.argmax(1), W - 1 - cols[:, ::-1].argmax(1),
                                            H - 1 - rows[:, ::-1].argmax(1)], axis=1).astype(np.int64))
        if self.crop:
            templates = self.rgb_proposal_processor(templates, bboxes)
        intrinsic = torch.tensor([[600, 0, 210], [0, 600, 210], [0, 0, 1]]).reshape(3, 3)
        return {"templates": templates, "masks": masks, "depths": depths, "model_name": model_name, "tar_file": tar_name,
                "intrinsic": intrinsic}
