"""Generates tests/golden/u_video_cut_golden.npz from the reference's golden output files.

Run in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Source files (README-documented output of `geotrax batch data/U_video_cut.mp4 --no-geo`, /root/reference/data/README.md:15-19):
    /root/reference/data/results-pixel/U_video_cut.txt             (19,817 x 14: frame,id,x,y,w,h,xs,ys,ws,hs,cls,conf,len,wid)
    /root/reference/data/results-pixel/U_video_cut_vid_transf.txt  (149 x 10: frame, H row-major)
The fixture keeps every transform and the track rows of a frame subset (enough to pin box-warp semantics);
u_video_cut_tracks_full.npz keeps the whole 19,817 x 14 table for the post-processing pin.
"""
import os
import numpy as np

REF = "/root/reference/data/results-pixel"
HERE = os.path.dirname(os.path.abspath(__file__))
FRAMES = [0, 1, 2, 3, 10, 25, 50, 75, 100, 125, 148, 149]

tracks = np.loadtxt(os.path.join(REF, "U_video_cut.txt"), delimiter=",")
transf = np.loadtxt(os.path.join(REF, "U_video_cut_vid_transf.txt"), delimiter=",")
keep = np.isin(tracks[:, 0].astype(int), FRAMES)
per_frame = np.bincount(tracks[:, 0].astype(int), minlength=150)
np.savez_compressed(os.path.join(HERE, "u_video_cut_golden.npz"),
                    tracks=tracks[keep], transforms=transf, dets_per_frame=per_frame,
                    n_rows_total=np.int64(len(tracks)))
print("rows kept", int(keep.sum()), "of", len(tracks), "; transforms", transf.shape)
# the whole track table too (14 columns = the OUTPUT of postprocess_tracks: the last two are estimate_vehicle_dimensions' length / width):
# tests/test_postprocess.py recomputes them from the first twelve with the vectorised mirror
np.savez_compressed(os.path.join(HERE, "u_video_cut_tracks_full.npz"), tracks=tracks)
