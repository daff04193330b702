"""Drop-in for the CLIP branch of the reference's `run_visual_tokenization.py` (function names, argument
meaning and the returned dict are the reference's; line numbers below refer to that file).

What changes underneath:
  * frames of MANY videos are encoded per native call (`frame_batch`, default 256) instead of one 8-frame
    video at a time (:226-254) — the tower is the tcgen05 path of vidil_clip_forward;
  * `image_embeds @ text_embeds.t()` (:276) and the per-frame host `np.argsort(...)[::-1][:k]` (:298-306) are one
    device call (vidil_sim_topk): the [F, T] score matrix never exists on the host, only [F, k] indices do;
  * the text bank is embedded by the text tower alone (the reference pushes DUMMY_IMAGE through the vision
    tower once per 512 phrases, :90-91, and discards the result).
Everything with ordering semantics — prompt prefixes (:57-80), the ontology clean-up including its
remove-while-iterating quirk (:386-396), frame_tokens layout (:268,300-308), frequency aggregation (:173-187),
the per-rank slice (:427-434) and the rank-ordered merge (:453-457) — is kept as is.
"""
from __future__ import annotations

import json
import os
from collections import defaultdict

import torch

from . import distributed as vdist
from . import jsonio
from . import ops

EMBBDING_BATCH_LIMIT_TEXT = 512  # :470 (spelling is the reference's)
OMIT_KEYWORDS = ['media player', 'video', 'playing video', 'audio', 'sound', 'taking video', 'water mark',
                 'water marked', 'watermark', 'watermarks', 'for sale in', 'sold from', 'stock', 'sold on', 'by viewers',
                 'are provided by', 'are posted on', 'for more', 'tag with', 'stream from', 'viewed from',
                 'showing video of', 'are on at', 'shuttlecock', 'shutter', 'shutter is white', 'shutters have bones',
                 'tape is looped', 'bliss wants you', 'thumbnail', 'technique']  # :471-472

ONTOLOGY_FILES = {  # :371-383, relative to the ontology root (the reference's working directory)
    'vg': dict(objects='visual_token_ontology/vg/openimage_classes_all_cleaned_fictional_characters.json',
               attributes='visual_token_ontology/vg/vg_original_attributes_synsets_keys_cleaned_remove_similar0.9.json',
               scenes='visual_token_ontology/vg/place365_ontology.json',
               verbs='visual_token_ontology/vg/vg_srl_selected_object_synsets_keys_remove_similar0.9.json'),
    'vg_tencent': dict(objects='visual_token_ontology/vg_tencent/tencent_ml_images_objects.json',
                       attributes='visual_token_ontology/vg_tencent/vg_original_attributes_synsets_keys_cleaned_remove_similar0.9.json',
                       scenes='visual_token_ontology/vg/place365_ontology.json',
                       verbs='visual_token_ontology/vg_tencent/vg_srl_selected_object_synsets_keys_remove_similar0.9.json'),
}


def load_json(json_path):
    with open(json_path) as f:
        return json.load(f)


def save_json(filepath, json_object):
    with open(filepath, 'w') as f:
        jsonio.dump_indent4(json_object, f)       # == json.dump(json_object, f, indent=4), the reference's line :52


def get_prefix_prompt_functions(version):
    """:57-80 — 'v0' identity, 'v1' 'A photo of {x}' for all four phrase types."""
    if version == 'v0':
        fn = lambda x: x  # noqa: E731
    elif version == 'v1':
        fn = lambda x: f'A photo of {x}'  # noqa: E731
    else:  # the reference falls through to an UnboundLocalError here
        raise UnboundLocalError(f"unknown prompt version {version!r}")
    return {'objects': fn, 'attributes': fn, 'scenes': fn, 'verbs': fn}


def load_ontology(ontology: str, root: str = '.') -> dict:
    """:368-406 — the four phrase lists after the reference's clean-up, quirk included: attributes that are also
    objects are removed from the list *while iterating over it* (:389-391), which skips the element after every
    removal; the surviving list is whatever that loop leaves, so the same loop runs here."""
    if ontology not in ONTOLOGY_FILES:
        raise KeyError(f"unknown ontology {ontology!r} (the reference knows 'vg' and 'vg_tencent')")
    paths = ONTOLOGY_FILES[ontology]
    object_texts = load_json(os.path.join(root, paths['objects']))
    attribute_texts = load_json(os.path.join(root, paths['attributes']))
    scene_texts = load_json(os.path.join(root, paths['scenes']))
    verb_texts = load_json(os.path.join(root, paths['verbs']))
    if isinstance(verb_texts, dict):
        verb_texts = list(verb_texts.keys())
    for key in attribute_texts:
        if key in object_texts:
            attribute_texts.remove(key)
    for key in OMIT_KEYWORDS:
        for lst in (object_texts, attribute_texts, scene_texts, verb_texts):
            if key in lst:
                lst.remove(key)
    return {'objects': object_texts, 'attributes': attribute_texts, 'scenes': scene_texts, 'verbs': verb_texts}


def _to_device(inputs, device):
    return inputs.to(device) if hasattr(inputs, "to") else {k: v.to(device) for k, v in inputs.items()}


@torch.no_grad()
def get_text_embeddings_clip(model, processor, texts, device, shard_over_ranks: bool = False):
    """:84-96 — [len(texts), proj] unit-norm text embeddings in batches of 512; returns (embeds, None, None).

    shard_over_ranks (extra): under torch.distributed every rank embeds a contiguous run of the 512-phrase batches and the
    rows are exchanged with one all-gather (vdist.all_gather_rows), instead of every rank embedding the whole bank as the
    reference does.  The batches are the same batches, each row depends on its own phrase only, and the concatenation is in
    the original order, so the bank is identical on every rank."""
    starts = list(range(0, len(texts), EMBBDING_BATCH_LIMIT_TEXT))
    if shard_over_ranks and vdist.get_world_size() > 1:
        per = (len(starts) + vdist.get_world_size() - 1) // vdist.get_world_size()
        starts = starts[vdist.get_rank() * per:(vdist.get_rank() + 1) * per]
    text_embeds = []
    for i in starts:
        text = texts[i: i + EMBBDING_BATCH_LIMIT_TEXT]
        inputs = _to_device(processor(text=text, return_tensors="pt", padding=True, truncation=True), device)
        outputs = model(input_ids=inputs["input_ids"], attention_mask=inputs.get("attention_mask"))
        text_embeds.append(outputs.text_embeds)
    if shard_over_ranks and vdist.get_world_size() > 1:
        mine = torch.cat(text_embeds, dim=0) if text_embeds else None
        width = torch.tensor([mine.shape[1] if mine is not None else 0], dtype=torch.int64, device=device)
        torch.distributed.all_reduce(width, op=torch.distributed.ReduceOp.MAX)     # a rank without batches still needs D
        if mine is None:
            mine = torch.zeros(0, int(width.item()), dtype=torch.float32, device=device)
        return vdist.all_gather_rows(mine.float()), None, None
    return torch.cat(text_embeds, dim=0), None, None


@torch.no_grad()
def get_image_embeddings_clip(model, processor, images, device):
    """:136-143 — images: list of PIL images (or anything the processor accepts) -> (None, image_embeds [n, proj])."""
    inputs = _to_device(processor(images=images, return_tensors="pt"), device)
    outputs = model(pixel_values=inputs["pixel_values"])
    return None, outputs.image_embeds


def aggregate_frame_tokens(frame_tokens):
    """:173-187 — phrases counted over frames in (rank j outer, frame i inner) insertion order, stable sort by count
    descending, first `topk` kept; a key whose first-frame list is empty stays []."""
    keys = frame_tokens[0].keys()
    aggregated_tokens = {key: [] for key in keys}
    topk = len(frame_tokens[0]['objects'])
    num_frm = len(frame_tokens)
    for key in keys:
        if frame_tokens[0][key] == []:
            continue
        count_dict = defaultdict(int)
        for j in range(topk):
            for i in range(num_frm):
                count_dict[frame_tokens[i][key][j]] += 1
        candidates = sorted(count_dict.items(), key=lambda x: x[1], reverse=True)
        aggregated_tokens[key] = [item[0] for item in candidates[:topk]]
    return aggregated_tokens


def tokens_from_embeddings(image_embeds: torch.Tensor, text_representations: dict, visual_token_texts: dict,
                           video_ids: list, captions: list, num_frm: int, topk_visualize: int) -> dict:
    """:268-312 for the CLIP branch: per bank one fused similarity + top-k on the device, then phrase lookup and
    aggregation on the host.  image_embeds [len(video_ids) * num_frm, proj] on the device."""
    assert image_embeds.shape[0] == len(video_ids) * num_frm, \
        "every kept video must contribute exactly num_frm frames (the reference's .view at :298 assumes it)"
    out = {vid: {"frame_tokens": [defaultdict(list) for _ in range(num_frm)], "caption": captions[i]}
           for i, vid in enumerate(video_ids)}
    for key, phrases in visual_token_texts.items():
        text_embeds = text_representations[key]['text_embeds']
        _, idx = ops.sim_topk(image_embeds, text_embeds, topk_visualize)
        idx = idx.view(len(video_ids), num_frm, -1).cpu().tolist()     # [V, num_frm, k] ints: the only D2H
        for j, vid in enumerate(video_ids):
            for frm_idx in range(num_frm):
                out[vid]['frame_tokens'][frm_idx][key] = [phrases[ii] for ii in idx[j][frm_idx]]
    for obj in out.values():
        obj["aggregated_tokens"] = aggregate_frame_tokens(obj['frame_tokens'])
    return out


@torch.no_grad()
def predict_video(config, video_dataset, model, device, visual_token_texts, prompt_functions, encoder_version='clip',
                  processor=None, frame_batch: int = 256, shard_text_bank: bool = False):
    """Same arguments and result as the reference's predict_video (:161-314) for encoder_version='clip':
    {video_id: {"frame_tokens": [{key: [k phrases]} x num_frm], "caption": ..., "aggregated_tokens": {key: [...]}}}."""
    if encoder_version != 'clip':
        raise NotImplementedError("the native path covers encoder_version='clip' (the pipeline default, "
                                  "run_frame_captioning_and_visual_tokenization.sh:23)")
    model.eval()
    num_frm = config['num_frm_visual_tokenization']
    text_representations = {}
    for key in visual_token_texts.keys():
        texts = [prompt_functions[key](t) for t in visual_token_texts[key]]
        text_embeds, text_ids, text_atts = get_text_embeddings_clip(model, processor, texts, device, shard_over_ranks=shard_text_bank)
        text_representations[key] = {'text_embeds': text_embeds, 'text_ids': text_ids, 'text_atts': text_atts}

    image_embeds, video_ids, captions = [], [], []
    pending = []

    def flush():
        if pending:
            image_embeds.append(get_image_embeddings_clip(model, processor, pending, device)[1])
            pending.clear()

    for i in range(len(video_dataset)):
        if i == config.get("early_stop_step", -1):
            print(f'early stop at {i}')
            break
        ann = video_dataset.annotation[i]
        video_name = os.path.basename(ann['video'])[:-4]
        frames, caption = video_dataset[i]
        if frames is None:  # :236-238 — unloadable videos are skipped, not fatal
            print('skip video that cannot be loaded:', video_name)
            continue
        pending.extend(list(frames))
        video_ids.append(video_name)
        captions.append(caption)
        if len(pending) >= frame_batch:
            flush()
    flush()
    if not video_ids:
        return {}
    image_embeds = torch.cat(image_embeds, dim=0)
    return tokens_from_embeddings(image_embeds, text_representations, visual_token_texts, video_ids, captions, num_frm,
                                  config['topk_visualize'])


def run(config, video_dataset, model, processor, device, visual_token_texts, output_dir: str | None = None,
        frame_batch: int = 256):
    """main() of the reference from the rank split onwards (:426-463): slice the annotation list for this rank,
    predict, and merge on rank 0 — through one all-gather of JSON rows instead of tmp/{rank}.json files."""
    start, end = vdist.shard_bounds(len(video_dataset.annotation))
    video_dataset.annotation = video_dataset.annotation[start:end]
    prompt_functions = get_prefix_prompt_functions(config['prompt_version_visual_tokenization'])
    result = predict_video(config, video_dataset, model, device, visual_token_texts, prompt_functions,
                           encoder_version='clip', processor=processor, frame_batch=frame_batch,
                           shard_text_bank=vdist.get_world_size() > 1)
    path = os.path.join(output_dir, 'visual_tokens.json') if output_dir else None
    return vdist.gather_and_write(result, path)
