"""Host side of `run_video_CapFilt.py` on the native path: caption every sampled frame of a video with the BLIP decoder,
de-duplicate, filter the candidates with the ITM head, merge the per-rank results.

The functions keep the reference's names, arguments and result (`caption_frames` :93-105, `filter_captions` :107-126,
`CapFilt` :139-204, the rank split and merge of `main` :237-291); the models they drive are the drop-ins of
`vidil_b200.blip`.  Three things differ, none visible in the output:

  * `filter_captions` scores every (caption, frame) pair of a video in ONE native call (`BLIP_ITM.forward_ids` with
    `frame_of_seq`) instead of one filterer call — and one ViT pass — per caption (:110-112);
  * video decoding (decord, :38-93) and sentence splitting (spaCy, :141,163-170) are not on the accelerated path: they come
    in as the callables `frame_loader` and `sentence_splitter`, defaulting to the reference's own libraries when installed;
  * the per-rank tmp files + barrier + rank-0 re-read (:258-291) become one all-gather of JSON rows.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from . import distributed as vdist
from . import jsonio
from .preprocess import process_frame, process_frames  # noqa: F401  (process_frame re-exported, :128-137)


@torch.no_grad()
def caption_frames(captioner, images, mode='beam'):
    """run_video_CapFilt.py:93-105."""
    if mode == 'beam':
        return captioner.generate(images, sample=False, num_beams=3, max_length=20, min_length=5)
    return captioner.generate(images, sample=True, top_p=0.9, max_length=20, min_length=5)


def _itm_prob(itm_logits: torch.Tensor) -> np.ndarray:
    return torch.nn.functional.softmax(itm_logits, dim=1)[:, 1].detach().cpu().numpy()     # :113


def _reduce_prob(itm_score: np.ndarray, mode: str) -> float:
    if mode == 'avg_filter':
        return np.sum(itm_score) / len(itm_score)                                          # :116-117
    if mode == 'max_filter':
        return np.max(itm_score)                                                           # :118-119
    raise UnboundLocalError(f"filter mode {mode!r}: the reference defines prob only for 'avg_filter' and 'max_filter'")


@torch.no_grad()
def filter_captions(filterer, images, texts, threshold, mode='max_filter'):
    """run_video_CapFilt.py:107-126: keep caption t iff max (or mean) over the frames of softmax(itm(frame, t))[1] > threshold.
    All len(texts) x len(images) pairs go through one native call."""
    if len(texts) == 0:
        return []
    n_frames = images.size()[0]
    tok = filterer.tokenizer
    if tok is None:
        raise RuntimeError("filter_captions needs filterer.tokenizer (bert-base-uncased is not on disk: pass tokenizer= to BLIP_ITM)")
    text = tok(list(texts), padding='max_length', truncation=True, max_length=35, return_tensors="pt")    # blip_itm.py:46-47
    # pairs in frame-major order (frame j, caption i) -> row j * len(texts) + i: the captions of a frame form one query group
    ids = text.input_ids.repeat(n_frames, 1)
    mask = text.attention_mask.repeat(n_frames, 1)
    logits = filterer.forward_ids(images, ids, mask, seqs_per_frame=len(texts))
    itm_score = _itm_prob(logits).reshape(n_frames, len(texts)).T                          # [caption, frame]
    filtered_captions = []
    for i, t in enumerate(texts):
        if _reduce_prob(itm_score[i], mode) > threshold:
            filtered_captions.append(t)
    return filtered_captions


def _default_frame_loader(video_path, strategy, num_frm):
    """load_video_from_path_decord (:38-93) needs decord, which is not part of this package."""
    raise RuntimeError("no frame_loader given and video decoding is outside vidil_b200 (the reference uses decord): pass "
                       "frame_loader(video_path, frm_sampling_strategy, num_frm) -> uint8 [num_frm, H, W, 3] or None")


def _default_sentence_splitter():
    import spacy
    nlp = spacy.load("en_core_web_sm", disable=['ner', 'tagger', 'lemmatizer'])                       # :141
    return lambda caption: [sent.text for sent in nlp(caption).sents]


def dedup_exact(captions):
    """:184-188 — first occurrence wins, order kept."""
    final = []
    for cap in captions:
        if cap not in final:
            final.append(cap)
    return final


@torch.no_grad()
def filter_captions_batched(filterer, images_list, texts_list, threshold, mode='max_filter'):
    """filter_captions for several videos in ONE native call: video v's captions texts_list[v] are scored against its own frames
    images_list[v] only (ragged: any number of captions per video).  Returns one filtered list per video, each identical to
    filter_captions(filterer, images_list[v], texts_list[v], threshold, mode)."""
    tok = filterer.tokenizer
    if tok is None:
        raise RuntimeError("filter_captions needs filterer.tokenizer (bert-base-uncased is not on disk: pass tokenizer= to BLIP_ITM)")
    flat_texts, seq_text, seq_frame, spans = [], [], [], []
    frame0 = 0
    for images, texts in zip(images_list, texts_list):
        n_frames = images.size()[0]
        spans.append((len(seq_text), len(texts), n_frames))
        for t in texts:                                           # caption-major inside a video, like filter_captions' loop
            seq_text += [len(flat_texts)] * n_frames
            seq_frame += list(range(frame0, frame0 + n_frames))
            flat_texts.append(t)
        frame0 += n_frames
    if not flat_texts:
        return [[] for _ in texts_list]
    text = tok(flat_texts, padding='max_length', truncation=True, max_length=35, return_tensors="pt")
    sel = torch.tensor(seq_text, dtype=torch.long)
    logits = filterer.forward_ids(torch.cat(list(images_list), dim=0), text.input_ids[sel], text.attention_mask[sel],
                                  frame_of_seq=torch.tensor(seq_frame, dtype=torch.int32))
    itm_score = _itm_prob(logits)
    out = []
    for (start, n_texts, n_frames), texts in zip(spans, texts_list):
        kept = []
        for i, t in enumerate(texts):
            if _reduce_prob(itm_score[start + i * n_frames:start + (i + 1) * n_frames], mode) > threshold:
                kept.append(t)
        out.append(kept)
    return out


@torch.no_grad()
def CapFilt(data, config, device, captioner=None, filterer=None, frame_loader=None, sentence_splitter=None, frame_processor=None,
            video_batch=1):
    """run_video_CapFilt.py:139-204; mutates the items of `data` exactly as the reference does ('unfiltered_text', 'text').
    frame_processor(frames_u8 [n,H,W,3] on device, image_size) -> float [n,3,S,S]; default: the native process_frames.
    video_batch > 1: that many videos share one ViT pass, one beam search and one ITM call (the reference does one video at a time
    with 4-8 frames, which leaves a B200 idle); every item ends up exactly as with video_batch=1, because a frame's caption and a
    (caption, frame) score do not depend on what else is in the batch."""
    from .blip import blip_decoder, blip_itm
    if config.get("caption") and captioner is None:
        captioner = blip_decoder(pretrained=config["caption_model_ckpt"], image_size=config["image_size"], vit=config["vit"])
    if captioner is not None:
        captioner = captioner.eval().to(device)
    if config.get("filter") and filterer is None:
        filterer = blip_itm(pretrained=config["filterer_model_ckpt"], image_size=config["image_size"], vit=config["vit"])
    if filterer is not None:
        filterer = filterer.eval().to(device)
    frame_loader = frame_loader or _default_frame_loader
    frame_processor = frame_processor or process_frames
    do_split = ('do_sentence_tokenization' not in config) or config['do_sentence_tokenization']
    if do_split and sentence_splitter is None:
        sentence_splitter = _default_sentence_splitter()

    data = list(data)
    for c0 in range(0, len(data), max(1, video_batch)):
        loaded = []                                              # (item, processed_frms)
        for item in data[c0:c0 + max(1, video_batch)]:
            video_path = item['video_path']
            try:
                raw_sample_frms = frame_loader(video_path, config["frm_sampling_strategy"], config["num_frm_CapFilt"])
                frames_u8 = torch.as_tensor(np.asarray(raw_sample_frms)).to(device)
                loaded.append((item, frame_processor(frames_u8, config["image_size"])))   # = stack(process_frame(f) ...), :161
            except Exception:  # noqa: BLE001 - the reference skips anything that fails to load (:162)
                print(f'skip video that cannot be loaded: {video_path}')
        if not loaded:
            continue

        # captioning (:172-194), one beam search for all frames of the chunk
        per_item_generated = [[] for _ in loaded]
        if config["caption"]:
            counts = [frms.size()[0] for _, frms in loaded]
            flat = caption_frames(captioner, torch.cat([frms for _, frms in loaded], dim=0), mode=config["generation_mode"])
            pos = 0
            for k, n in enumerate(counts):
                per_item_generated[k] = dedup_exact(flat[pos:pos + n])                     # :184-188
                pos += n

        to_filter = []
        for k, (item, _) in enumerate(loaded):
            if do_split:
                original_caption_sentences = []
                for original_cap in item['text']:
                    original_caption = original_cap.replace('\n', '. ')
                    for sent in sentence_splitter(original_caption):
                        if len(sent) > 3:
                            original_caption_sentences.append(sent.strip())
            else:
                original_caption_sentences = [cap.replace('\n', '. ').strip() for cap in item['text']]
            generated_captions_final = per_item_generated[k]
            if not config["caption"]:
                candidate_captions = original_caption_sentences
                item['unfiltered_text'] = candidate_captions
            elif config['keep_original_caption']:
                candidate_captions = original_caption_sentences + generated_captions_final
                item['unfiltered_text'] = candidate_captions
            else:
                item['text'] = []
                candidate_captions = generated_captions_final
                item['unfiltered_text'] = candidate_captions
            if config["filter"]:
                to_filter.append(generated_captions_final if config["filter_generated_only"] else candidate_captions)
            else:
                item['text'] = candidate_captions

        # filtering (:196-203), one ITM call for every (caption, frame) pair of the chunk
        if config["filter"]:
            if len(loaded) == 1:
                kept = [filter_captions(filterer, loaded[0][1], to_filter[0], config["threshold"], config['filter_mode'])]
            else:
                kept = filter_captions_batched(filterer, [frms for _, frms in loaded], to_filter, config["threshold"],
                                               config['filter_mode'])
            for (item, _), texts in zip(loaded, kept):
                if config["filter_generated_only"]:
                    item['text'] += texts
                else:
                    item['text'] = texts


def collect_rank_outputs(items):
    """:250-259 — the two dicts a rank contributes."""
    filtered, unfiltered = {}, {}
    for item in items:
        if 'unfiltered_text' not in item:
            print(f"skip video that cannot be loaded: {item['video_path']}")
            continue
        unfiltered[item['video_id']] = item['unfiltered_text']
        if item['text'] != []:
            filtered[item['video_id']] = item['text']
        else:
            print('filter out video:', item['video_id'])
    return filtered, unfiltered


def run(data, config, device, output_dir=None, **capfilt_kwargs):
    """main() from the rank split on (:237-291): this rank's contiguous slice -> CapFilt -> merged dicts on rank 0, written as
    video_text_CapFilt.json / video_text_Cap.json with the reference's json.dump(indent=4)."""
    start, end = vdist.shard_bounds(len(data))
    mine = data[start:end]
    CapFilt(mine, config, device, **capfilt_kwargs)
    filtered, unfiltered = collect_rank_outputs(mine)
    per_rank = vdist.all_gather_json({"filtered": filtered, "unfiltered": unfiltered})
    if not vdist.is_main_process():
        return None
    merged_f = vdist.merge_rank_dicts([r["filtered"] if r else None for r in per_rank])
    merged_u = vdist.merge_rank_dicts([r["unfiltered"] if r else None for r in per_rank])
    if output_dir is not None:
        os.makedirs(output_dir, exist_ok=True)
        with open(os.path.join(output_dir, 'video_text_CapFilt.json'), 'w') as out:
            jsonio.dump_indent4(merged_f, out)    # == json.dump(..., indent=4), run_video_CapFilt.py:283-291
        with open(os.path.join(output_dir, 'video_text_Cap.json'), 'w') as out:
            jsonio.dump_indent4(merged_u, out)
    return merged_f, merged_u
