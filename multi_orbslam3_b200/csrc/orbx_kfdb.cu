// orbx_kfdb.cu - the server's keyframe-descriptor database resident on one GPU (SURVEY 8f row 3 -> 8e).
//
// A keyframe reaches the server as a KF.msg (R/msg/KF.msg:24-29): `CvKeyPoint[] mvKeysUn` = N records of 15 packed bytes
// (R/msg/CvKeyPoint.msg, written by Converter::toCvKeyPointMsg, R/src/Converter.cc:218-230) and `Descriptor[] mDescriptors`
// = N x uint8[32] (R/msg/Descriptor.msg:1; a fixed-size array carries no length prefix, so the N descriptors are 32 N
// contiguous bytes of the serialised message).  The reference rebuilds a cv::Mat and a std::vector<cv::KeyPoint> from them
// element by element (KeyFrame.cc:1929-1944, Converter.cc:232-244).  Here the two byte runs go to the device as they are:
// the descriptor run is appended to this GPU's shard of the descriptor DB by one asynchronous copy (no host-side repack),
// the keypoint records are unpacked by k_kp_from_msg on the device.  The shard is what orbx_bf_knn2_device and
// multi_orbslam3_b200.server.ShardedDescriptorDB search.
#include <algorithm>
#include <mutex>
#include "orbx_match_internal.h"

void orbx_launch_kp_from_msg(const uint8_t* d_msg15, int n, orbx_keypoint* d_kps, cudaStream_t s);   // orbx_search.cu

struct KfEntry { int64_t id; long long first; int n; };

struct orbx_kfdb {
    int device;
    long long cap_rows, rows;
    int max_kf;
    uint8_t* d_desc;            // [cap_rows][32]
    orbx_keypoint* d_kps;       // [cap_rows]
    uint8_t* d_msg;             // staging of the keypoint wire records of the keyframes in flight: ring of STAGE_ROWS x 15 bytes
    long long stage_rows, stage_head;
    cudaStream_t stream;
    std::vector<KfEntry> kf;    // in ingestion order: first rows ascending
    std::mutex mu;              // the communication threads of several clients may ingest concurrently (Communicator.cc:110-148)
};

extern "C" int orbx_kfdb_create(int device, long long capacity_rows, int max_keyframes, orbx_kfdb** out)
{
    if (!out || capacity_rows <= 0 || capacity_rows > 0x7fffffffLL || max_keyframes <= 0) return ORBX_E_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        orbx_set_error("%s: no such CUDA device (%s)", "orbx_kfdb_create", "there is no CPU fallback");
        return ORBX_E_CUDA;
    }
    CKM(cudaSetDevice(device));
    orbx_kfdb* db = new orbx_kfdb();
    db->device = device; db->cap_rows = capacity_rows; db->rows = 0; db->max_kf = max_keyframes;
    db->stage_rows = std::min<long long>(capacity_rows, 1 << 20); db->stage_head = 0;
    db->d_desc = nullptr; db->d_kps = nullptr; db->d_msg = nullptr; db->stream = nullptr;
    cudaError_t e = cudaMalloc(&db->d_desc, (size_t)capacity_rows * 32);
    if (e == cudaSuccess) e = cudaMalloc(&db->d_kps, (size_t)capacity_rows * sizeof(orbx_keypoint));
    if (e == cudaSuccess) e = cudaMalloc(&db->d_msg, (size_t)db->stage_rows * 15);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&db->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        orbx_set_error("%s failed: %s", "orbx_kfdb_create", cudaGetErrorString(e));
        cudaFree(db->d_desc); cudaFree(db->d_kps); cudaFree(db->d_msg);
        delete db;
        return e == cudaErrorMemoryAllocation ? ORBX_E_NOMEM : ORBX_E_CUDA;
    }
    db->kf.reserve(max_keyframes);
    *out = db;
    return ORBX_OK;
}

extern "C" void orbx_kfdb_destroy(orbx_kfdb* db)
{
    if (!db) return;
    cudaSetDevice(db->device);
    cudaStreamSynchronize(db->stream);
    cudaFree(db->d_desc); cudaFree(db->d_kps); cudaFree(db->d_msg);
    cudaStreamDestroy(db->stream);
    delete db;
}

// reserves rows [first, first + n) and a keyframe entry; the caller holds the lock
static int reserve(orbx_kfdb* db, int64_t kf_id, int n, long long* first)
{
    if (db->rows + n > db->cap_rows || (int)db->kf.size() >= db->max_kf) {
        orbx_set_error("%s: the keyframe DB is full (%s)", "orbx_kfdb", db->rows + n > db->cap_rows ? "rows" : "keyframes");
        return ORBX_E_CAPACITY;
    }
    *first = db->rows;
    db->kf.push_back(KfEntry{kf_id, db->rows, n});
    db->rows += n;
    return ORBX_OK;
}

extern "C" int orbx_kfdb_ingest_msg(orbx_kfdb* db, int64_t kf_id, const uint8_t* msg_keys15, const uint8_t* msg_desc32, int n,
                                    long long* first_row)
{
    if (!db || n < 0 || (n > 0 && !msg_desc32)) return ORBX_E_INVALID;
    std::lock_guard<std::mutex> lk(db->mu);
    CKM(cudaSetDevice(db->device));
    long long first = 0;
    int rc = reserve(db, kf_id, n, &first);
    if (rc) return rc;
    if (first_row) *first_row = first;
    if (n == 0) return ORBX_OK;
    CKM(cudaMemcpyAsync(db->d_desc + (size_t)first * 32, msg_desc32, (size_t)n * 32, cudaMemcpyHostToDevice, db->stream));
    if (msg_keys15) {
        // wire records through the staging ring (15-byte records cannot be unpacked in place: the result is 28 bytes)
        for (int done = 0; done < n;) {
            if (db->stage_head == db->stage_rows) db->stage_head = 0;      // ring wrapped: the stream orders the copy after the earlier unpack
            const int take = (int)std::min<long long>(n - done, db->stage_rows - db->stage_head);
            uint8_t* dm = db->d_msg + (size_t)db->stage_head * 15;
            CKM(cudaMemcpyAsync(dm, msg_keys15 + (size_t)done * 15, (size_t)take * 15, cudaMemcpyHostToDevice, db->stream));
            orbx_launch_kp_from_msg(dm, take, db->d_kps + first + done, db->stream);
            CKM(cudaGetLastError());
            db->stage_head += take; done += take;
        }
    } else {
        CKM(cudaMemsetAsync(db->d_kps + first, 0, (size_t)n * sizeof(orbx_keypoint), db->stream));
    }
    return ORBX_OK;
}

extern "C" int orbx_kfdb_ingest_slot_device(orbx_kfdb* db, int64_t kf_id, orbx_extractor* ex, int slot, long long* first_row)
{
    if (!db || !ex) return ORBX_E_INVALID;
    if (orbx_ex_device(ex) != db->device) { orbx_set_error("%s: extractor and DB live on different devices%s", "orbx_kfdb_ingest_slot_device", ""); return ORBX_E_INVALID; }
    orbx_keypoint* dk; uint8_t* dd; int32_t* dn; int cap, slots;
    int rc = orbx_extractor_results_device(ex, &dk, &dd, &dn, nullptr, &cap, &slots);
    if (rc) return rc;
    if (slot < 0 || slot >= slots) return ORBX_E_INVALID;
    rc = orbx_extractor_sync(ex, nullptr);             // the slot must be complete (and free of deferred errors)
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(db->mu);
    CKM(cudaSetDevice(db->device));
    int32_t n = 0;
    CKM(cudaMemcpy(&n, dn + slot, sizeof(n), cudaMemcpyDeviceToHost));
    if (n < 0 || n > cap) return ORBX_E_INVALID;
    long long first = 0;
    rc = reserve(db, kf_id, n, &first);
    if (rc) return rc;
    if (first_row) *first_row = first;
    if (n == 0) return ORBX_OK;
    CKM(cudaMemcpyAsync(db->d_desc + (size_t)first * 32, dd + (size_t)slot * cap * 32, (size_t)n * 32, cudaMemcpyDeviceToDevice, db->stream));
    CKM(cudaMemcpyAsync(db->d_kps + first, dk + (size_t)slot * cap, (size_t)n * sizeof(orbx_keypoint), cudaMemcpyDeviceToDevice, db->stream));
    return ORBX_OK;
}

extern "C" int orbx_kfdb_size(orbx_kfdb* db, long long* rows, int* keyframes, long long* capacity_rows)
{
    if (!db) return ORBX_E_INVALID;
    std::lock_guard<std::mutex> lk(db->mu);
    if (rows) *rows = db->rows;
    if (keyframes) *keyframes = (int)db->kf.size();
    if (capacity_rows) *capacity_rows = db->cap_rows;
    return ORBX_OK;
}

extern "C" int orbx_kfdb_device(orbx_kfdb* db, const uint8_t** d_desc, const orbx_keypoint** d_kps)
{
    if (!db) return ORBX_E_INVALID;
    if (d_desc) *d_desc = db->d_desc;
    if (d_kps) *d_kps = db->d_kps;
    return ORBX_OK;
}

extern "C" int orbx_kfdb_sync(orbx_kfdb* db)
{
    if (!db) return ORBX_E_INVALID;
    CKM(cudaSetDevice(db->device));
    CKM(cudaStreamSynchronize(db->stream));
    return ORBX_OK;
}

extern "C" int orbx_kfdb_locate(orbx_kfdb* db, const long long* rows, int n, int64_t* kf_id, int32_t* feature)
{
    if (!db || n < 0 || (n > 0 && (!rows || !kf_id || !feature))) return ORBX_E_INVALID;
    std::lock_guard<std::mutex> lk(db->mu);
    for (int i = 0; i < n; i++) {
        const long long r = rows[i];
        kf_id[i] = -1; feature[i] = -1;
        if (r < 0 || r >= db->rows) continue;
        // last keyframe whose first row is <= r (empty keyframes share a first row with their successor: skip them)
        auto it = std::upper_bound(db->kf.begin(), db->kf.end(), r, [](long long v, const KfEntry& e) { return v < e.first; });
        --it;
        kf_id[i] = it->id; feature[i] = (int32_t)(r - it->first);
    }
    return ORBX_OK;
}

extern "C" int orbx_kfdb_knn2(orbx_kfdb* db, orbx_matcher* m, const uint8_t* q, int nq, long long idx_base, int32_t* idx, int32_t* dist)
{
    if (!db || !m || nq < 0 || (nq > 0 && (!q || !idx || !dist))) return ORBX_E_INVALID;
    if (m->p.device != db->device) { orbx_set_error("%s: matcher and DB live on different devices%s", "orbx_kfdb_knn2", ""); return ORBX_E_INVALID; }
    if (nq == 0) return ORBX_OK;
    long long rows;
    { std::lock_guard<std::mutex> lk(db->mu); rows = db->rows; }
    if (idx_base < 0 || idx_base + rows > 0x7fffffffLL) return ORBX_E_INVALID;
    CKM(cudaSetDevice(db->device));
    CKM(cudaStreamSynchronize(db->stream));            // everything ingested so far is searchable
    int rc = orbx_m_gen_scratch(m, (size_t)nq * (32 + 16));
    if (rc) return rc;
    uint8_t* dq = m->d_gen; int32_t* di = reinterpret_cast<int32_t*>(dq + (size_t)nq * 32); int32_t* dd = di + (size_t)nq * 2;
    cudaStream_t s = m->stream;
    CKM(cudaMemcpyAsync(dq, q, (size_t)nq * 32, cudaMemcpyHostToDevice, s));
    rc = orbx_bf_knn2_device(m, dq, nq, db->d_desc, rows, di, dd, (int)idx_base, s);
    if (rc) return rc;
    CKM(cudaMemcpyAsync(idx, di, sizeof(int32_t) * 2 * nq, cudaMemcpyDeviceToHost, s));
    CKM(cudaMemcpyAsync(dist, dd, sizeof(int32_t) * 2 * nq, cudaMemcpyDeviceToHost, s));
    CKM(cudaStreamSynchronize(s));
    return ORBX_OK;
}
