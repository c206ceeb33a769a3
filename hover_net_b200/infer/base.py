"""Drop-in for reference `infer/base.py:22-94` (`InferManager`): same constructor arguments, same
`run_step` / `post_proc_func` attributes, same type-colour table and JSON writer -- with the model
and the post-processing bound to libhvn instead of torch + cv2/scipy/skimage.

Differences that follow from the engine (documented, not silent): no `torch.nn.DataParallel`
(one process per GPU; set LOCAL_RANK / HVN_DEVICE), and a checkpoint may also be given as a dict /
`.npz` of arrays besides the reference's `torch.load(path)["desc"]` file."""
import json
import os

import numpy as np


class InferManager(object):
    def __init__(self, **kwargs):
        self.run_step = None
        self.type_info_path = None
        for variable, value in kwargs.items():
            self.__setattr__(variable, value)
        self.__load_model()
        self.nr_types = self.method["model_args"]["nr_types"]
        # create type info name and colour (reference infer/base.py:31-53)
        self.type_info_dict = {None: ["no label", [0, 0, 0]]}
        if self.nr_types is not None and self.type_info_path is not None:
            self.type_info_dict = json.load(open(self.type_info_path, "r"))
            self.type_info_dict = {int(k): (v[0], tuple(v[1])) for k, v in self.type_info_dict.items()}
            for k in range(self.nr_types):
                if k not in self.type_info_dict:
                    assert False, "Not detect type_id=%d defined in json." % k
        if self.nr_types is not None and self.type_info_path is None:
            # the reference samples matplotlib's "hot" colormap at integer positions 0..nr_types-1, i.e.
            # its first entries (near-black reds); the same values without the matplotlib dependency
            # its first entries: red ramps linearly from 0.0416 at x = 0 to 1.0 at x = 0.365079 (x = k / 255), green and
            # blue stay 0 there; `(cmap(k)[:3] * 255).astype(uint8)` truncates (reference infer/base.py:46-48)
            hot = [(int((0.0416 + (1.0 - 0.0416) * (k / 255.0) / 0.365079) * 255), 0, 0) for k in range(self.nr_types)]
            self.type_info_dict = {k: (str(k), hot[k]) for k in range(self.nr_types)}
        return

    def __load_model(self):
        """Create the model, load the checkpoint and bind the two plugin functions
        (reference infer/base.py:56-78)."""
        from ..models.hovernet import net_desc, post_proc, run_desc

        net = net_desc.create_model(**self.method["model_args"])
        src = self.method["model_path"]
        if isinstance(src, dict):
            state = src
        elif str(src).endswith(".npz"):
            state = dict(np.load(src))
        else:
            import torch
            state = torch.load(src, map_location="cpu")["desc"]
        net.load_state_dict(state, strict=True)  # strips an optional `module.` prefix itself
        self.net = net
        self.device_tile_path = True  # infer/tile.py: pad / patch / stitch / crop / process on the device
        self.run_step = lambda input_batch: run_desc.infer_step(input_batch, net)
        self.post_proc_func = post_proc.process
        return

    def _save_json(self, path, old_dict, mag=None):
        """reference infer/base.py:80-94 (format contract: SURVEY.md App. D)."""
        new_dict = {}
        for inst_id, inst_info in old_dict.items():
            new_inst_info = {}
            for info_name, info_value in inst_info.items():
                if isinstance(info_value, np.ndarray):
                    info_value = info_value.tolist()
                new_inst_info[info_name] = info_value
            new_dict[int(inst_id)] = new_inst_info
        json_dict = {"mag": mag, "nuc": new_dict}
        with open(path, "w") as handle:
            json.dump(json_dict, handle)
        return new_dict

    # the reference's subclasses call self.__save_json, which name-mangles to this
    _InferManager__save_json = _save_json
