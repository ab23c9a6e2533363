"""Model-directory resolution shared by every ``from_pretrained`` / ``load_*_weights`` entry.

A local directory is used as is; anything else is treated as a HuggingFace Hub repo id and fetched with a lazily
imported ``huggingface_hub`` exactly where the reference does it (model.py:170-176, encoder.py:342-344,
decoder.py:274-276: ``snapshot_download``; config.py:139-148: ``hf_hub_download`` of ``config.json``).  Offline the hub
call raises its own error, as it would in the reference."""
from __future__ import annotations

import json
from pathlib import Path


def model_dir(model_id_or_path, **kwargs) -> Path:
    path = Path(model_id_or_path)
    if path.is_dir():
        return path
    from huggingface_hub import snapshot_download

    return Path(snapshot_download(repo_id=str(model_id_or_path), **kwargs))


def config_dict(model_id_or_path) -> dict:
    path = Path(model_id_or_path)
    if path.is_dir():
        config_file = path / "config.json"
    else:
        from huggingface_hub import hf_hub_download

        config_file = Path(hf_hub_download(repo_id=str(model_id_or_path), filename="config.json"))
    return json.loads(Path(config_file).read_text(encoding="utf-8"))
