/* TEST INFRASTRUCTURE.  The reference's FeatureAssociation tears its thread pool down with flags = 0
 * (LSD/myFA.cpp:63, "known bug: a task may not finish"), which DROPS tasks still queued: the set of scored
 * hypotheses then varies from run to run.  oracle/_ref/libref_glibc_det.so links the unmodified sources with
 * -Wl,--wrap=threadpool_destroy so that the pool drains its queue first (threadpool_graceful): the reference's
 * arithmetic, made deterministic, as the comparator for the drop-in FeatureAssociation test. */
typedef struct threadpool_t threadpool_t;
int __real_threadpool_destroy(threadpool_t* pool, int flags);
int __wrap_threadpool_destroy(threadpool_t* pool, int flags) {
    (void)flags;
    return __real_threadpool_destroy(pool, 1 /* threadpool_graceful, LSD/threadpool.h:69 */);
}
