/* orc_mq.h -- MQ coder state shared by orc_mq.c and orc_t1.c (oracle, test infrastructure only). */
#ifndef ORC_MQ_H
#define ORC_MQ_H
#include <stdint.h>

/* context ids, internal/entropy/mqc.go:135-166 */
enum {
    ORC_CTX_ZC0 = 0, ORC_CTX_SC0 = 9, ORC_CTX_MAG0 = 14, ORC_CTX_RL = 17, ORC_CTX_UNI = 18,
    ORC_NUM_CTX = 19
};

/* 94-entry tables (2*state+mps), mqc.go:21-132 */
extern uint32_t orc_mq_qe[94];
extern uint8_t  orc_mq_nmps[94];
extern uint8_t  orc_mq_nlps[94];
void orc_mq_tables_init(void);

typedef struct {
    uint32_t A, C, CT;
    const uint8_t *data; int len; int bp;
    uint8_t ctx[ORC_NUM_CTX];
} orc_mqdec;

void orc_mqdec_init(orc_mqdec *d, const uint8_t *data, int len);
int  orc_mqdec_decode(orc_mqdec *d, int ctx);
#endif
