// NVDEC ingest: H.264 / HEVC elementary stream -> NV12 frames in HBM, without the frames ever crossing PCIe uncompressed.
//
// Replaces the reference's reader (`cv2.VideoCapture(...).read()`: FFmpeg software decode + swscale to BGR24 on the host,
// /root/reference/geotrax/extract.py:146, 248) for callers that hold the bitstream: SURVEY.md section 8f rank 1 / 8a-2.  The decoded NV12
// surfaces feed gt_preprocess / gt_extract_batch directly (gt_set_input_format(GT_INPUT_NV12): fused NV12 letterbox kernel).
//
// libnvcuvid.so.1 is part of the driver, not of the CUDA toolkit, and this image has no nvcuvid.h: the library is dlopen()ed and the few
// structures the API needs are declared here (Video Codec SDK `cuviddec.h` / `nvcuvid.h` layout, stable since SDK 9: fields this file does not use are
// reserved space).  Without the library every entry point returns GT_ERR_STATE and gt_nvdec_available() is 0 -- nothing falls back to a
// CPU decoder.
#include <dlfcn.h>

#include <deque>

#include "engine.cuh"

namespace {

typedef void* CUvideoparser;
typedef void* CUvideodecoder;
typedef void* CUvideoctxlock;
typedef long long CUvideotimestamp;

enum { cudaVideoCodec_H264 = 4, cudaVideoCodec_HEVC = 8 };
enum { cudaVideoSurfaceFormat_NV12 = 0 };
enum { cudaVideoDeinterlaceMode_Weave = 0 };
enum { cudaVideoChromaFormat_420 = 1 };
enum { cudaVideoCreate_PreferCUVID = 4 };
enum { CUVID_PKT_ENDOFSTREAM = 0x01 };

struct CUVIDDECODECREATEINFO {
  unsigned long ulWidth, ulHeight, ulNumDecodeSurfaces;
  int CodecType, ChromaFormat;
  unsigned long ulCreationFlags, bitDepthMinus8, ulIntraDecodeOnly, ulMaxWidth, ulMaxHeight, Reserved1;
  struct { short left, top, right, bottom; } display_area;
  int OutputFormat, DeinterlaceMode;
  unsigned long ulTargetWidth, ulTargetHeight, ulNumOutputSurfaces;
  CUvideoctxlock vidLock;
  struct { short left, top, right, bottom; } target_rect;
  unsigned long enableHistogram;
  unsigned long Reserved2[4];
};

struct CUVIDEOFORMAT {
  int codec;
  struct { unsigned int numerator, denominator; } frame_rate;
  unsigned char progressive_sequence, bit_depth_luma_minus8, bit_depth_chroma_minus8, min_num_decode_surfaces;
  unsigned int coded_width, coded_height;
  struct { int left, top, right, bottom; } display_area;
  int chroma_format;
  unsigned int bitrate;
  struct { int x, y; } display_aspect_ratio;
  struct { unsigned char flags, color_primaries, transfer_characteristics, matrix_coefficients; } video_signal_description;
  unsigned int seqhdr_data_length;
};

struct CUVIDPARSERDISPINFO {
  int picture_index, progressive_frame, top_field_first, repeat_first_field;
  CUvideotimestamp timestamp;
};

struct CUVIDPICPARAMS;   // opaque here: handed from the parser to cuvidDecodePicture unchanged

typedef int (*PFNVIDSEQUENCECALLBACK)(void*, CUVIDEOFORMAT*);
typedef int (*PFNVIDDECODECALLBACK)(void*, CUVIDPICPARAMS*);
typedef int (*PFNVIDDISPLAYCALLBACK)(void*, CUVIDPARSERDISPINFO*);

struct CUVIDPARSERPARAMS {
  int CodecType;
  unsigned int ulMaxNumDecodeSurfaces, ulClockRate, ulErrorThreshold, ulMaxDisplayDelay;
  unsigned int bAnnexb_and_reserved;
  unsigned int uReserved1[4];
  void* pUserData;
  PFNVIDSEQUENCECALLBACK pfnSequenceCallback;
  PFNVIDDECODECALLBACK pfnDecodePicture;
  PFNVIDDISPLAYCALLBACK pfnDisplayPicture;
  void* pfnGetOperatingPoint;
  void* pfnGetSEIMsg;
  void* pvReserved2[5];
  void* pExtVideoInfo;
};

struct CUVIDSOURCEDATAPACKET {
  unsigned long flags, payload_size;
  const unsigned char* payload;
  CUvideotimestamp timestamp;
};

struct CUVIDPROCPARAMS {
  int progressive_frame, second_field, top_field_first, unpaired_field;
  unsigned int reserved_flags, reserved_zero;
  unsigned long long raw_input_dptr;
  unsigned int raw_input_pitch, raw_input_format;
  unsigned long long raw_output_dptr;
  unsigned int raw_output_pitch, Reserved1;
  CUstream output_stream;
  unsigned int Reserved[46];
  unsigned long long* histogram_dptr;
  void* Reserved2[1];
  unsigned char tail_guard[256];   // head-room should a newer driver's structure be longer
};

struct Api {
  void* lib = nullptr;
  CUresult (*CreateVideoParser)(CUvideoparser*, CUVIDPARSERPARAMS*) = nullptr;
  CUresult (*ParseVideoData)(CUvideoparser, CUVIDSOURCEDATAPACKET*) = nullptr;
  CUresult (*DestroyVideoParser)(CUvideoparser) = nullptr;
  CUresult (*CreateDecoder)(CUvideodecoder*, CUVIDDECODECREATEINFO*) = nullptr;
  CUresult (*DestroyDecoder)(CUvideodecoder) = nullptr;
  CUresult (*DecodePicture)(CUvideodecoder, CUVIDPICPARAMS*) = nullptr;
  CUresult (*MapVideoFrame64)(CUvideodecoder, int, unsigned long long*, unsigned int*, CUVIDPROCPARAMS*) = nullptr;
  CUresult (*UnmapVideoFrame64)(CUvideodecoder, unsigned long long) = nullptr;
  bool ok = false;
};

Api& api() {
  static Api a;
  static bool tried = false;
  if (tried) return a;
  tried = true;
  a.lib = dlopen("libnvcuvid.so.1", RTLD_NOW | RTLD_LOCAL);
  if (!a.lib) return a;
#define SYM(field, name) *(void**)(&a.field) = dlsym(a.lib, name)
  SYM(CreateVideoParser, "cuvidCreateVideoParser"); SYM(ParseVideoData, "cuvidParseVideoData"); SYM(DestroyVideoParser, "cuvidDestroyVideoParser");
  SYM(CreateDecoder, "cuvidCreateDecoder"); SYM(DestroyDecoder, "cuvidDestroyDecoder"); SYM(DecodePicture, "cuvidDecodePicture");
  SYM(MapVideoFrame64, "cuvidMapVideoFrame64"); SYM(UnmapVideoFrame64, "cuvidUnmapVideoFrame64");
#undef SYM
  a.ok = a.CreateVideoParser && a.ParseVideoData && a.DestroyVideoParser && a.CreateDecoder && a.DestroyDecoder && a.DecodePicture &&
         a.MapVideoFrame64 && a.UnmapVideoFrame64;
  return a;
}

}  // namespace

struct gt_decoder {
  gt_engine* e = nullptr;
  int codec = 0;
  CUvideoparser parser = nullptr;
  CUvideodecoder decoder = nullptr;
  cudaStream_t stream = nullptr;
  int width = 0, height = 0;           // display size = the engine's frame size
  uint8_t* ring = nullptr;             // [cap][height * 3 / 2][width] dense NV12
  int cap = 0, head = 0, count = 0;    // frames [head, head + count) (mod cap) are decoded and not yet taken
  size_t frame_bytes = 0;
  int error = 0;                       // sticky error raised inside a parser callback
  std::string err;
};

namespace {

int fail(gt_decoder* d, const char* what, int code) {
  char buf[256];
  snprintf(buf, sizeof(buf), "nvdec: %s (code %d)", what, code);
  d->err = buf;
  d->error = 1;
  return 0;   // a parser callback returning 0 aborts cuvidParseVideoData
}

int on_sequence(void* user, CUVIDEOFORMAT* f) {
  gt_decoder* d = (gt_decoder*)user;
  const int w = f->display_area.right - f->display_area.left, h = f->display_area.bottom - f->display_area.top;
  if (f->chroma_format != cudaVideoChromaFormat_420 || f->bit_depth_luma_minus8 != 0) return fail(d, "only 8-bit 4:2:0 streams are supported", f->chroma_format);
  if (w != d->width || h != d->height) return fail(d, "stream frame size differs from the engine's frame_w x frame_h", w);
  const int nsurf = std::max(4, (int)f->min_num_decode_surfaces + 2);
  if (d->decoder) return nsurf;   // same geometry: keep the decoder
  CUVIDDECODECREATEINFO ci;
  memset(&ci, 0, sizeof(ci));
  ci.ulWidth = f->coded_width; ci.ulHeight = f->coded_height; ci.ulNumDecodeSurfaces = (unsigned long)nsurf;
  ci.CodecType = f->codec; ci.ChromaFormat = f->chroma_format;
  ci.ulCreationFlags = cudaVideoCreate_PreferCUVID;
  ci.bitDepthMinus8 = 0;
  ci.ulMaxWidth = f->coded_width; ci.ulMaxHeight = f->coded_height;
  ci.display_area.left = (short)f->display_area.left; ci.display_area.top = (short)f->display_area.top;
  ci.display_area.right = (short)f->display_area.right; ci.display_area.bottom = (short)f->display_area.bottom;
  ci.OutputFormat = cudaVideoSurfaceFormat_NV12; ci.DeinterlaceMode = cudaVideoDeinterlaceMode_Weave;
  ci.ulTargetWidth = (unsigned long)w; ci.ulTargetHeight = (unsigned long)h;
  ci.ulNumOutputSurfaces = 2;
  const CUresult r = api().CreateDecoder(&d->decoder, &ci);
  if (r != CUDA_SUCCESS) return fail(d, "cuvidCreateDecoder failed (no NVDEC engine visible in this container?)", (int)r);
  return nsurf;
}

int on_decode(void* user, CUVIDPICPARAMS* pp) {
  gt_decoder* d = (gt_decoder*)user;
  if (!d->decoder) return fail(d, "picture before sequence header", 0);
  const CUresult r = api().DecodePicture(d->decoder, pp);
  if (r != CUDA_SUCCESS) return fail(d, "cuvidDecodePicture failed", (int)r);
  return 1;
}

// display order: map the surface, copy its two planes into the next dense ring slot, unmap
int on_display(void* user, CUVIDPARSERDISPINFO* di) {
  gt_decoder* d = (gt_decoder*)user;
  if (!di) return 1;   // end of stream
  if (d->count >= d->cap) return fail(d, "decoded-frame ring is full: take frames (gt_decoder_take) or feed smaller chunks", d->cap);
  CUVIDPROCPARAMS vp;
  memset(&vp, 0, sizeof(vp));
  vp.progressive_frame = di->progressive_frame; vp.top_field_first = di->top_field_first; vp.unpaired_field = di->repeat_first_field < 0;
  vp.output_stream = (CUstream)d->stream;
  unsigned long long src = 0;
  unsigned int pitch = 0;
  CUresult r = api().MapVideoFrame64(d->decoder, di->picture_index, &src, &pitch, &vp);
  if (r != CUDA_SUCCESS) return fail(d, "cuvidMapVideoFrame failed", (int)r);
  uint8_t* dst = d->ring + (size_t)((d->head + d->count) % d->cap) * d->frame_bytes;
  // the mapped surface is pitch-linear NV12: luma rows, then (at row `height` aligned as the driver laid it out = height rows) interleaved chroma
  cudaError_t ce = cudaMemcpy2DAsync(dst, (size_t)d->width, (const void*)src, pitch, (size_t)d->width, (size_t)d->height, cudaMemcpyDeviceToDevice, d->stream);
  if (ce == cudaSuccess)
    ce = cudaMemcpy2DAsync(dst + (size_t)d->width * d->height, (size_t)d->width, (const void*)(src + (unsigned long long)pitch * d->height), pitch, (size_t)d->width,
                           (size_t)d->height / 2, cudaMemcpyDeviceToDevice, d->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(d->stream);   // the surface goes back to the decoder at unmap
  api().UnmapVideoFrame64(d->decoder, src);
  if (ce != cudaSuccess) return fail(d, cudaGetErrorString(ce), (int)ce);
  d->count++;
  return 1;
}

}  // namespace

extern "C" {

int gt_nvdec_available(void) { return api().ok ? 1 : 0; }

const char* gt_decoder_last_error(gt_decoder_handle d) { return d ? d->err.c_str() : "null decoder"; }

int gt_decoder_create(gt_handle e, int codec, int capacity_frames, gt_decoder_handle* out) {
  if (!e || !out) return GT_ERR_INVALID;
  *out = nullptr;
  GT_CHECK(e, codec == GT_CODEC_H264 || codec == GT_CODEC_HEVC, "gt_decoder_create: unknown codec %d", codec);
  GT_CHECK(e, api().ok, "gt_decoder_create: libnvcuvid.so.1 (NVDEC driver library) is not available on this machine");
  GT_CHECK(e, (e->cfg.frame_w % 2) == 0 && (e->cfg.frame_h % 2) == 0, "gt_decoder_create: NV12 needs even frame dimensions");
  if (cudaSetDevice(e->device) != cudaSuccess) { gt_set_error(e, "cudaSetDevice failed"); return GT_ERR_CUDA; }
  cudaFree(0);   // the primary context is current on this thread: cuvid* use it
  gt_decoder* d = new gt_decoder();
  d->e = e; d->codec = codec; d->width = e->cfg.frame_w; d->height = e->cfg.frame_h;
  d->frame_bytes = (size_t)d->width * d->height * 3 / 2;
  d->cap = std::max(capacity_frames, 2 * e->cfg.max_batch);
  if (cudaMalloc((void**)&d->ring, (size_t)d->cap * d->frame_bytes) != cudaSuccess || cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking) != cudaSuccess) {
    gt_set_error(e, "gt_decoder_create: out of device memory for %d NV12 frames", d->cap);
    if (d->ring) cudaFree(d->ring);
    delete d;
    return GT_ERR_NOMEM;
  }
  CUVIDPARSERPARAMS pp;
  memset(&pp, 0, sizeof(pp));
  pp.CodecType = codec == GT_CODEC_HEVC ? cudaVideoCodec_HEVC : cudaVideoCodec_H264;
  pp.ulMaxNumDecodeSurfaces = 1;      // the sequence callback returns the real number
  pp.ulMaxDisplayDelay = 0;           // low latency: pictures are displayed as soon as they are decodable in order
  pp.pUserData = d;
  pp.pfnSequenceCallback = on_sequence; pp.pfnDecodePicture = on_decode; pp.pfnDisplayPicture = on_display;
  const CUresult r = api().CreateVideoParser(&d->parser, &pp);
  if (r != CUDA_SUCCESS) {
    gt_set_error(e, "cuvidCreateVideoParser failed: %d", (int)r);
    cudaFree(d->ring); cudaStreamDestroy(d->stream);
    delete d;
    return GT_ERR_CUDA;
  }
  *out = d;
  return GT_OK;
}

int gt_decoder_destroy(gt_decoder_handle d) {
  if (!d) return GT_OK;
  cudaSetDevice(d->e->device);
  if (d->parser) api().DestroyVideoParser(d->parser);
  if (d->decoder) api().DestroyDecoder(d->decoder);
  if (d->stream) cudaStreamDestroy(d->stream);
  if (d->ring) cudaFree(d->ring);
  delete d;
  return GT_OK;
}

int gt_decoder_feed(gt_decoder_handle d, const uint8_t* data, size_t size) {
  if (!d) return GT_ERR_INVALID;
  if (cudaSetDevice(d->e->device) != cudaSuccess) return GT_ERR_CUDA;
  CUVIDSOURCEDATAPACKET pkt;
  memset(&pkt, 0, sizeof(pkt));
  pkt.payload = data; pkt.payload_size = (unsigned long)size;
  if (!data || size == 0) pkt.flags = CUVID_PKT_ENDOFSTREAM;
  d->error = 0;
  const CUresult r = api().ParseVideoData(d->parser, &pkt);
  if (d->error) { gt_set_error(d->e, "%s", d->err.c_str()); return GT_ERR_STATE; }
  if (r != CUDA_SUCCESS) { gt_set_error(d->e, "cuvidParseVideoData failed: %d", (int)r); d->err = d->e->err; return GT_ERR_CUDA; }
  return GT_OK;
}

int gt_decoder_pending(gt_decoder_handle d) { return d ? d->count : GT_ERR_INVALID; }

int gt_decoder_take(gt_decoder_handle d, int max_frames, const uint8_t** dev_nv12, int32_t* n_frames) {
  if (!d || !dev_nv12 || !n_frames || max_frames < 1) return GT_ERR_INVALID;
  const int n = std::min(std::min(max_frames, d->count), d->cap - d->head);   // a contiguous run of the ring
  *dev_nv12 = d->ring + (size_t)d->head * d->frame_bytes;
  *n_frames = n;
  d->head = (d->head + n) % d->cap;
  d->count -= n;
  return GT_OK;
}

}  // extern "C"
