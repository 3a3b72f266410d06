"""Replay of the 2D memory-bank validation step of the reference driver (`func_2d/function.py:423-534`, SURVEY §3.4 /
§8(f) rank 4) against ANY model object with the reference's submodule surface — the real reference (golden generator),
or the product (tests).  The driver's own glue (bank stacking, cosine-similarity sampling, permutes) is repeated here
in plain torch exactly as the driver writes it; the multinomial draw is replaced by a recorded index tensor so that the
run is deterministic.  Inputs are seeded; `sub`-sampling keeps the fixture small."""
import torch
import torch.nn.functional as F


def make_inputs(B_img, n_bank, n_prompts, size=512, seed=11):
    g = torch.Generator().manual_seed(seed)
    e = size // 16
    img = torch.randn(B_img, 3, size, size, generator=g)
    bank = [(torch.randn(1, 64, e, e, generator=g), torch.randn(1, 64, e, e, generator=g), 0.5,
             torch.randn(256 * e * e, generator=g)) for _ in range(n_bank)]
    pts = torch.rand(n_prompts, 1, 2, generator=g) * size
    labels = torch.ones(n_prompts, 1, dtype=torch.int32)
    hi_mask = (torch.rand(B_img, 1, size, size, generator=g) > 0.6).float()
    return img, bank, pts, labels, hi_mask


def replay(net, img, bank, pts, labels, hi_mask, sampled_indices=None, device="cpu"):
    """-> dict of tensors: similarity, sampled_indices, memattn, low_res, iou, obj, maskmem_feat, maskmem_pos"""
    out = {}
    img = img.to(device)
    e = img.shape[-1] // 16
    feat_sizes = [(4 * e, 4 * e), (2 * e, 2 * e), (e, e)]
    backbone_out = net.forward_image(img)                                                   # function.py:423
    _, vision_feats, vision_pos_embeds, _ = net._prepare_backbone_features(backbone_out)    # :424
    vision_feats, vision_pos_embeds = list(vision_feats), list(vision_pos_embeds)
    B = vision_feats[-1].size(1)
    to_cat_memory, to_cat_memory_pos, to_cat_image_embed = [], [], []
    for element in bank:                                                                    # :434-439
        to_cat_memory.append(element[0].to(device).flatten(2).permute(2, 0, 1))
        to_cat_memory_pos.append(element[1].to(device).flatten(2).permute(2, 0, 1))
        to_cat_image_embed.append(element[3].to(device))
    memory_stack_ori = torch.stack(to_cat_memory, dim=0)
    memory_pos_stack_ori = torch.stack(to_cat_memory_pos, dim=0)
    image_embed_stack_ori = torch.stack(to_cat_image_embed, dim=0)
    vision_feats_temp = vision_feats[-1].float().permute(1, 0, 2).reshape(B, -1, e, e).reshape(B, -1)   # :446-447
    image_embed_stack_ori = F.normalize(image_embed_stack_ori, p=2, dim=1)
    vision_feats_temp = F.normalize(vision_feats_temp, p=2, dim=1)
    similarity_scores = torch.mm(image_embed_stack_ori, vision_feats_temp.t()).t()
    similarity_scores = F.softmax(similarity_scores, dim=1)                                 # :453
    out["similarity"] = similarity_scores
    if sampled_indices is None:                                                             # :454 (seeded here)
        sampled_indices = torch.multinomial(similarity_scores.cpu(), num_samples=B, replacement=True,
                                            generator=torch.Generator().manual_seed(5)).squeeze(1)
    sampled_indices = sampled_indices.to(device)
    out["sampled_indices"] = sampled_indices
    memory_stack_ori_new = memory_stack_ori[sampled_indices].squeeze(3).permute(1, 2, 0, 3)  # :456-460
    memory = memory_stack_ori_new.reshape(-1, memory_stack_ori_new.size(2), memory_stack_ori_new.size(3))
    memory_pos_stack_new = memory_pos_stack_ori[sampled_indices].squeeze(3).permute(1, 2, 0, 3)
    memory_pos = memory_pos_stack_new.reshape(-1, memory_stack_ori_new.size(2), memory_stack_ori_new.size(3))
    vision_feats[-1] = net.memory_attention(curr=[vision_feats[-1]], curr_pos=[vision_pos_embeds[-1]],   # :464-470
                                            memory=memory, memory_pos=memory_pos, num_obj_ptr_tokens=0)
    out["memattn"] = vision_feats[-1]
    feats = [feat.permute(1, 2, 0).reshape(B, -1, *fs) for feat, fs in zip(vision_feats[::-1], feat_sizes[::-1])][::-1]
    image_embed, high_res_feats = feats[-1], feats[:-1]                                     # :472-476
    se, de = net.sam_prompt_encoder(points=(pts.to(device), labels.to(device)), boxes=None, masks=None,
                                    batch_size=B)                                           # :483-488
    low_res, iou_pred, sam_tokens, obj = net.sam_mask_decoder(                              # :490-499
        image_embeddings=image_embed, image_pe=net.sam_prompt_encoder.get_dense_pe(), sparse_prompt_embeddings=se,
        dense_prompt_embeddings=de, multimask_output=False, repeat_image=False,
        cell_nums=torch.as_tensor([pts.shape[0]]).to(device), high_res_features=high_res_feats)
    out["low_res"], out["iou"], out["obj"] = low_res, iou_pred, obj
    mf, mp = net._encode_new_memory(current_vision_feats=vision_feats, feat_sizes=feat_sizes,            # :520-526
                                    pred_masks_high_res=hi_mask.to(device), is_mask_from_pts=True)
    out["maskmem_feat"], out["maskmem_pos"] = mf, mp[0]
    return out
