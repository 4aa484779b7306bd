// rg_engine_decode.inl -- entry points that need the MP3 decode kernels (stage A).
extern "C" int rg_analyze_tracks(rg_engine*, const rg_input*, size_t, rg_track_result*, uint32_t*) { return RG_ERR_UNSUPPORTED_FORMAT; }
extern "C" int rg_analyze_album(rg_engine*, const rg_input*, size_t, rg_track_result*, rg_album_result*, uint64_t*) { return RG_ERR_UNSUPPORTED_FORMAT; }
extern "C" int rg_find_peak(rg_engine*, const rg_input*, size_t, rg_peak_result*) { return RG_ERR_UNSUPPORTED_FORMAT; }
extern "C" int rg_corpus_upload(rg_engine*, const rg_input*, size_t, rg_corpus**) { return RG_ERR_UNSUPPORTED_FORMAT; }
extern "C" void rg_corpus_free(rg_engine*, rg_corpus* c) { delete c; }
extern "C" uint64_t rg_corpus_bytes(const rg_corpus* c) { return c ? c->total : 0; }
extern "C" int rg_analyze_corpus(rg_engine*, rg_corpus*, rg_track_result*, uint32_t*) { return RG_ERR_UNSUPPORTED_FORMAT; }
extern "C" int rg_analyze_album_shard(rg_engine*, rg_corpus*, rg_track_result*, uint64_t*, double*) { return RG_ERR_UNSUPPORTED_FORMAT; }
extern "C" int rg_album_finish(rg_engine*, const uint64_t*, const double*, rg_album_result*) { return RG_ERR_UNSUPPORTED_FORMAT; }
extern "C" int rg_decode_pcm(rg_engine*, const rg_input*, float*, size_t, size_t*, uint32_t*, uint32_t*) { return RG_ERR_UNSUPPORTED_FORMAT; }
